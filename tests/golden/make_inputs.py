"""Regenerate tests/inputs/*.txt from the reference's own test inputs.

Run in the build container only (needs /root/reference).  Each reference input is
parsed with the host mirror and re-emitted in normalised form, so the GPU box --
which has no /root/reference -- runs exactly the reference's configurations.
Settings the reference derives from defaults are written out explicitly.
Inputs that need GDAL / NetCDF (SRTM example, netcdf_restart) are skipped.
"""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kestrel_b200.host.inputfile import read_input_file, write_input_file  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "inputs")


def main():
    os.makedirs(OUT, exist_ok=True)
    files = sorted(glob.glob(os.path.join(REF, "tests", "Input_*.txt"))) + [os.path.join(REF, "examples", "Input1d_cap_constslope.txt")]
    for f in files:
        name = os.path.basename(f)
        if "netcdf" in name:
            continue
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rs = read_input_file(f)
        stem = name.replace("Input_", "").replace("Input", "").replace(".txt", "")
        rel = os.path.relpath(f, REF)
        write_input_file(rs, os.path.join(OUT, f"case_{stem}.txt"),
                         header=f"normalised from jakelangham/kestrel {rel} by tests/golden/make_inputs.py")
        print("wrote", stem)


if __name__ == "__main__":
    main()
