"""Writes tests/golden/reference_state_dict_keys.txt: name, shape of every entry of the reference
PixelNeRF.state_dict() under the shipped config (configs/train_dtu.yaml:31-50).  Build container only."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from oracle import ref_import  # noqa: E402

ns = ref_import.load()
sys.modules["diner_ref_image_encoder"] = ns.image_encoder
sys.modules["diner_ref_resnetfc"] = ns.resnetfc
D = ref_import._DotMap
m = ns.pixelnerf.PixelNeRF(
    poscode_conf=D(kwargs=dict(num_freqs=6, freq_factor=6.28, include_input=True)),
    encoder_conf=D(module="diner_ref_image_encoder.SpatialEncoder", kwargs=dict(image_padding=64, padding_pe=4, pretrained=False)),
    mlp_fine_conf=D(module="diner_ref_resnetfc.ResnetFC", kwargs=dict(n_blocks=5, d_hidden=512, combine_layer=3, combine_type="average")))
with open(os.path.join(ROOT, "tests", "golden", "reference_state_dict_keys.txt"), "w") as f:
    for k, v in m.state_dict().items():
        f.write("%s %s\n" % (k, "x".join(map(str, v.shape)) or "scalar"))
print(len(m.state_dict()))
