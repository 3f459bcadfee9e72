"""The `opt` fields and head description the hot path reads (reference
sgtapose/lib/opts_parallel.py:31,107-142,392-395,429-430).  The reference's
argparse namespace can be passed instead of this one unchanged."""
import types

HEADS = {"hm": 7, "reg": 2, "wh": 2, "tracking": 2}
HEAD_CONV = {h: [256] for h in HEADS}
ARCH = "dlapawdl3new_34"


def default_opt(**over):
    o = types.SimpleNamespace(
        pre_img=True, pre_hm=True, ct_modify=False, head_kernel=3, prior_bias=-4.6,
        dla_node="dcn", load_model="x", model_output_list=False, num_classes=7,
        pos_embed=True, zero_tracking=False,
        k_list_1=1, k_list_2=1, k_list_3=1, k_list_4=1, k_list_5=1, k_list_6=1,
        ks1=12, ks2=6, ks3=3, ks4=1, ks5=1, ks6=1)
    for k, v in over.items():
        setattr(o, k, v)
    return o
