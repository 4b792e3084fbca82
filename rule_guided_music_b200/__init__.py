"""B200-native (sm_100a) implementation of the rule-guided SCG sampling hot path of yjhuangcd/rule-guided-music.

Sub-packages mirror the reference's module names for this path (guided_diffusion, music_rule_guidance, taming,
diff_collage); the compute lives in librgm_b200.so (hand-written CUDA, C ABI in include/rgm_b200.h).
"""
__version__ = "0.1.0"
