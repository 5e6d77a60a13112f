"""Compile the reference's own model code for the CPU-baseline / --impl reference legs of bench.py (TEST INFRASTRUCTURE).

The reference for this path is Python: `dgl/model/models.py` (+ `blocks.py`, its relative import).  The analogue of "compile the
reference from its sources where they lie, outputs only into oracle/_ref/" is byte-compilation: the files under
/root/reference/dgl/model/ are compiled IN PLACE and the marshalled code objects are written to oracle/_ref/model/*.bin
(git-ignored, but shipped to the GPU box like our own built .so; plain .pyc files are dropped by the gpurun snapshot, hence the
own container: one line with the interpreter version, then marshal.dumps(code)).  No reference source text enters the
repository, and /root/reference is not needed at run time.
Called by __graft_entry__.build() when /root/reference is present; a no-op elsewhere (the GPU box uses the prebuilt files)."""
import os
import marshal
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/dgl/model"
REF_OUT = os.path.join(HERE, "_ref", "model")
FILES = ("models.py", "blocks.py")
# the training script and the modules it imports (dgl/training.py:11-15): compiled the same way so that its unmodified train_dgl /
# test_dgl loops can be run over the compat packages on the GPU box (tests/test_compat.py)
SCRIPT_SRC = "/root/reference/dgl"
SCRIPT_OUT = os.path.join(HERE, "_ref", "scripts")
SCRIPT_FILES = ("training.py", "dataloader.py", "dataloader_utils.py", "utils.py")


def build_ref(force=False):
    """-> list of compiled files written (empty when the reference tree is absent)."""
    if not all(os.path.exists(os.path.join(REF_SRC, f)) for f in FILES):
        return []
    os.makedirs(REF_OUT, exist_ok=True)
    out = []
    os.makedirs(SCRIPT_OUT, exist_ok=True)
    todo = [(os.path.join(REF_SRC, f), os.path.join(REF_OUT, f[:-3] + ".bin")) for f in FILES]
    todo += [(os.path.join(SCRIPT_SRC, f), os.path.join(SCRIPT_OUT, f[:-3] + ".bin")) for f in SCRIPT_FILES if os.path.exists(os.path.join(SCRIPT_SRC, f))]
    for src, dst in todo:
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            with open(src, "rb") as fh, warnings.catch_warnings():
                warnings.simplefilter("ignore")          # the reference's own SyntaxWarnings (dgl/utils.py:47) are not ours to fix
                code = compile(fh.read(), src, "exec", dont_inherit=True)
            with open(dst, "wb") as fh:
                fh.write((sys.version.split()[0] + "\n").encode())
                fh.write(marshal.dumps(code))
        out.append(dst)
    return out


if __name__ == "__main__":
    print(build_ref(force=True))
