#!/usr/bin/env python
"""Record keys + shapes of the REFERENCE IASSD_Backbone.state_dict() (KITTI cfg) by instantiating the
reference's own class from oracle/_ref (installed by oracle/build_ref.sh).  CPU only.
    python tests/golden/make_state_dict_keys.py   ->  tests/golden/iassd_backbone_state_dict_keys.json"""
import importlib
import json
import sys
import warnings
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
from spsnet_b200 import backbone as bb  # noqa: E402

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    mod = importlib.import_module("pcdet.models.backbones_3d.IASSD_backbone")
ref = mod.IASSD_Backbone(bb.kitti_iassd_cfg(), num_class=3, input_channels=4)
out = {k: list(v.shape) for k, v in ref.state_dict().items()}
(Path(__file__).parent / "iassd_backbone_state_dict_keys.json").write_text(json.dumps(out, indent=0))
print(len(out), "entries")
