"""Builds `oracle/_ref/`: the UNMODIFIED reference's hot-path modules, byte-compiled from where they lie.

    python oracle/build_ref.py          (run by __graft_entry__.build() when /root/reference exists)

The reference is a directory of Python scripts (no setup.py, no native code), so "compiling the reference" means
`py_compile`: SCT-GAN/model.py (the model), SCT-GAN/train.py (its loss classes: SoliditySyntaxLoss,
ContractLevelFocalLoss, SpatialAwareFocalLoss) and SCT-GAN/data_augmentation.py (imported by train.py) are compiled
to `oracle/_ref/*.refbin` (pyc bytes under a neutral suffix: `*.pyc` files did not survive the snapshot to the GPU
box) — compiled artefacts only, no reference source enters the repository.  `oracle/_ref/` is
git-ignored (it stays out of history) but not gpurun-ignored, so it travels to the GPU box like our own .so, where
/root/reference does not exist.  `oracle/ref_loader.py` imports the three modules from those files.

This is test / measurement infrastructure: only tests/, __graft_entry__ and bench.py's CPU legs
(`cpu_baseline`, `--impl reference`) load it; the product path never does.
"""
import os
import py_compile
import sys

REF = "/root/reference/SCT-GAN"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = ("model", "train", "data_augmentation")


def build(force=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:
        src, dst = os.path.join(REF, name + ".py"), os.path.join(OUT, name + ".refbin")
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            # unchecked pyc: the loader must not look for the (absent) source file's mtime on the GPU box
            py_compile.compile(src, cfile=dst, dfile=f"reference/SCT-GAN/{name}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write(sys.version.split()[0] + "\n")
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else f"{REF} not present: nothing built")
