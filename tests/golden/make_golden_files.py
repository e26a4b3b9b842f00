#!/usr/bin/env python3
"""File-format fixtures: the reference's OWN on-disk artefacts for medium_data, produced by running the
unmodified reference writers in the build container (same scratch recipe as make_golden.py), so that the importers in
genie_smem_b200/surface.py are tested against what the reference really writes:

  ref_medium_data-FM.json.gz    ExactMatch.create_fm_index            (ExactMatch.py:22-33)
  ref_medium_data-LUT.json.gz   LUT.generate_lut(6) + save_lut        (LUT.py:15-35, 50-55)
  ref_rmi_medium_k6.pkl         RMI_LUT([10,100], 6).train_RMI + save (RMI_LUT.py:36-50, 186-190): sklearn models inside

Usage:  python tests/golden/make_golden_files.py [--scratch /tmp/genie_ref_scratch]
"""
import argparse
import gzip
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

OUT = mg.OUT


def gz_copy(src, name):
    with open(src, "rb") as f, gzip.open(os.path.join(OUT, name), "wb", compresslevel=9) as g:
        shutil.copyfileobj(f, g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scratch", default="/tmp/genie_ref_scratch")
    a = ap.parse_args()
    smem_dir = mg.setup(a.scratch)
    import ExactMatch as RE
    import LUT as RL
    import RMI_LUT as RR
    for stale in ("medium_data-FM.json", "medium_data-LUT.json"):
        p = os.path.join(smem_dir, "data", stale)
        if os.path.exists(p):
            os.remove(p)
    m = RE.ExactMatch("medium_data.fa")
    m.create_fm_index()
    gz_copy(os.path.join(smem_dir, "data", "medium_data-FM.json"), "ref_medium_data-FM.json.gz")
    lut = RL.LUT(m)
    lut.generate_lut(6)
    lut.save_lut()
    gz_copy(os.path.join(smem_dir, "data", "medium_data-LUT.json"), "ref_medium_data-LUT.json.gz")
    r = RR.RMI_LUT([10, 100], 6, "medium_data.fa")
    r.train_RMI()
    r.save("rmi_medium_k6.pkl")
    shutil.copy(os.path.join(smem_dir, "rmi_medium_k6.pkl"), os.path.join(OUT, "ref_rmi_medium_k6.pkl"))
    for f in ("ref_medium_data-FM.json.gz", "ref_medium_data-LUT.json.gz", "ref_rmi_medium_k6.pkl"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
