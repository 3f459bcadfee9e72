"""Shim package the reference imports at sgtapose/lib/model/networks/dla.py:21-25 (`from .DCNv2.dcn_v2 import DCN`).

Copy this directory to `sgtapose/lib/model/networks/DCNv2/` of a reference checkout (level 1) instead
of cloning and building lbin/DCNv2: `DCN` is then the B200 operator, with upstream's constructor keywords and
state-dict names.  No reference file changes."""
from sgtapose_b200.dcn_v2 import DCN, dcn_v2_conv  # noqa: F401
