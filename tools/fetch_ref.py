"""Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, but it travels to the GPU box).

    python tools/fetch_ref.py [--src /root/reference]

Nothing under baseline/_ref is product code or committed: it is the caller side of the drop-in
boundary (train_fns.py, utils/, cr_diff_aug.py, mycleanfid/) and the reference's own model files, used
  * by tests/test_gpu_dropin.py, which runs the reference's train_fns.GAN_training_function on the
    B200 modules of iea_gan_b200/dropin without touching a line of it, and
  * by `bench.py --impl reference`, which times the reference's own CPU implementation.
Files are copied byte for byte (sha256 listed in baseline/_ref/MANIFEST.json).  The modules the
reference imports at top level but this image lacks (boost_histogram, matplotlib, seaborn, cleanfid) are
not on the path of any function called here; empty stand-ins for them are written to
baseline/_ref/_stubs/ (caveat 1).
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = ["train_fns.py", "train.py", "cr_diff_aug.py", "config.json", "model.py", "layers.py", "RRM.py",
         "diff_aug.py", "loss.py", "LICENSE"]
DIRS = ["utils", "mycleanfid"]

STUBS = {
    "boost_histogram/__init__.py": "",
    "matplotlib/__init__.py": "from . import pyplot\n",
    "matplotlib/pyplot.py": "",
    "seaborn/__init__.py": "",
    "cleanfid/__init__.py": "",
    "cleanfid/downloads_helper.py": "",
    "cleanfid/inception_pytorch.py": "class InceptionV3(object):\n    pass\n",
    "cleanfid/resize.py": "",
    "cleanfid/utils.py": "",
    "cleanfid/features.py": "",
    "cleanfid/inception_torchscript.py": "",
}


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def fetch(src="/root/reference", quiet=False):
    if not os.path.isdir(src):
        raise FileNotFoundError("reference tree %s not present (the GPU box uses the staged copy)" % src)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(DST, f))
        manifest[f] = sha(os.path.join(DST, f))
    for d in DIRS:
        for base, _, names in os.walk(os.path.join(src, d)):
            for n in names:
                if not n.endswith(".py"):
                    continue
                rel = os.path.relpath(os.path.join(base, n), src)
                os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                shutil.copyfile(os.path.join(src, rel), os.path.join(DST, rel))
                manifest[rel] = sha(os.path.join(DST, rel))
    for rel, text in STUBS.items():
        p = os.path.join(DST, "_stubs", rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(text)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1)
    if not quiet:
        print("staged %d reference files under %s" % (len(manifest), DST))
    return DST


def present():
    return os.path.exists(os.path.join(DST, "train_fns.py"))


def activate(dropin):
    """Put the staged reference on sys.path.  dropin=True: the B200 modules answer to the names model /
    layers / RRM / diff_aug / loss and everything else (train_fns, utils, ...) is the reference's;
    dropin=False: the whole reference, for the CPU arm."""
    if not present():
        raise FileNotFoundError("baseline/_ref is not staged: run python tools/fetch_ref.py in the build container")
    for m in ("model", "layers", "RRM", "diff_aug", "loss", "train_fns", "utils", "cr_diff_aug", "mycleanfid"):
        for k in [k for k in sys.modules if k == m or k.startswith(m + ".")]:
            del sys.modules[k]
    paths = [os.path.join(DST, "_stubs"), DST]
    if dropin:
        paths.insert(0, os.path.join(ROOT, "iea_gan_b200", "dropin"))
    sys.path[:] = paths + [p for p in sys.path if p not in paths]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    fetch(ap.parse_args().src)
