"""Stages the UNMODIFIED reference for use as the CPU arm of the benchmark and as a live checker in the tests.

TEST / MEASUREMENT INFRASTRUCTURE, not product code: only tests/, __graft_entry__.smoke()/build() and bench.py's CPU
legs (`--impl reference`, `cpu_baseline`) may touch anything under oracle/.

The reference (Natasha-R/Flood-Prediction-GAN) is plain Python: nothing is compiled. This recipe copies its
`models/` package and the `metadata/dataset_split.csv` that `Model.__init__` reads (models/data.py:87) from
/root/reference -- which exists only in the authoring container -- into the git-ignored `oracle/_ref/`, byte for byte,
so that the copy travels to the GPU box with the repository snapshot (`.gitignore` lists oracle/_ref/, `.gpurunignore`
does not). No reference source enters the git history.

    python oracle/build_ref.py        # idempotent; prints what it staged
"""
import hashlib
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/model.py", "models/model_architectures.py", "models/data.py", "models/utils.py", "models/segmentation_model.py",
         "metadata/dataset_split.csv", "LICENSE"]


def stage(verbose=True):
    """Returns True when oracle/_ref holds the reference (freshly staged or already there), False when the reference
    is absent here and nothing was staged before."""
    if not os.path.isdir(REF):
        return os.path.isfile(os.path.join(DST, "models", "model.py"))
    manifest = []
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            if rel == "models/__init__.py":  # the reference's models/ is a namespace package
                continue
            raise FileNotFoundError(src)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append((rel, hashlib.sha256(open(dst, "rb").read()).hexdigest()[:16]))
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write("unmodified files of /root/reference staged by oracle/build_ref.py (sha256/16)\n")
        for rel, digest in manifest:
            f.write(f"{digest}  {rel}\n")
    if verbose:
        print(f"staged {len(manifest)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
