"""sgtapose_b200: B200-native (sm_100a) hot path of SGTAPose's per-frame dense inference.

Drop-in surfaces (reference file:line in each module's docstring):
  dcn_v2.DCN                  -- the DCNv2 operator module (dla.py:21-25, :545)
  networks.create_model       -- model factory / forward signature (model.py:24-29)
  decode.dream_generic_decode -- heatmap decode (decode.py:184-313)
"""
__version__ = "0.1.0"
