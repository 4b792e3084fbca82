"""Decoder half of the taming KL-VAE used by the sampling path (reference taming/models/klvae_pedal.py)."""
