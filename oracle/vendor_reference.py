"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- recipe that places the UNMODIFIED reference where the GPU box can run it.

    python -m oracle.vendor_reference        # (also called by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (seven modules + two .mat inputs, nothing to compile): the files are copied byte for byte from
/root/reference into oracle/_ref/ (git-ignored, so no reference source enters the history; not gpurun-ignored, so the
directory travels to the GPU box like the built .so).  `bench.py --impl reference` then times the reference's own
one_defl_Hutch_step there (cpu_baseline.kind = "reference"); oracle/ref_shim.py loads it from either place."""
import filecmp
import os
import shutil

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("examples.py", "gateway.py", "main.py", "matrix.py", "multigrid.py", "stoch_trace.py", "utils.py",
         "schwinger16.mat", "schwinger128.mat")


def vendor():
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if not os.path.isfile(d) or not filecmp.cmp(s, d, shallow=False):
            shutil.copyfile(s, d)
    return True


if __name__ == "__main__":
    print("vendored" if vendor() else "no reference tree at " + SRC, "->", DST)
