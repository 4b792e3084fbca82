"""Host-side mirror of the reference's guided_diffusion package for the SCG sampling path (B200-native backend)."""
