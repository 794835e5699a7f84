"""Copies the parity fixtures from the reference checkout into tests/golden/.

Inputs are the Middlebury pairs the reference's pics.txt names (data, not source); goldens
are the reference's own committed outputs.  /root/reference does
not exist on the GPU box, so the files are committed; this script documents where each one
came from and regenerates the directory.
"""
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/stereo_matching"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FILES = {
    "tsukuba": ["im1.png", "im5.png"], "teddy": ["im2.png", "im6.png"], "cones": ["im2.png", "im6.png"],
    "art": ["view1.png", "view5.png"], "laundry": ["view1.png", "view5.png"],
}
GOLDENS = ["asw_consistency_pre-reff.png", "asw_disparity.png", "asw_consistency_post-reff.png", "cross_based_initial.png",
           "cross_based_disparity.png"]
for ds, inputs in FILES.items():
    os.makedirs(os.path.join(OUT, ds), exist_ok=True)
    for f in inputs + GOLDENS:
        shutil.copyfile(os.path.join(REF, ds, f), os.path.join(OUT, ds, f))
os.makedirs(os.path.join(OUT, "sukub"), exist_ok=True)
for f in ["imL.png", "imP.png", "asw_raw_d.png"]:
    shutil.copyfile(os.path.join(REF, "sukub", f), os.path.join(OUT, "sukub", f))
shutil.copyfile(os.path.join(REF, "pics.txt"), os.path.join(OUT, "pics.txt"))
print("fixtures written to", OUT)
