"""Synthetic state dicts for the oracle's tests -- re-exported from the product package's generator
(rule_guided_music_b200/synthetic_weights.py: pure torch-CPU, no CUDA), so that the oracle, the golden-vector script
and the CUDA path are all fed bit-identical tensors.  The oracle package itself stays test infrastructure."""
from rule_guided_music_b200.synthetic_weights import (DIT_PRESETS, VAE_DDCONFIG, make_classifier_state_dict,  # noqa: F401
                                                       make_dit_state_dict, make_vae_encoder_state_dict, make_vae_state_dict,
                                                       vae_decoder_layout, vae_encoder_layout)
