"""Recipe that stages the reference's own hot-path modules for the CPU reference leg of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product package ``stardis_b200``.

The reference (tardis-sn/stardis) is pure Python + numba: there is nothing to compile.  ``/root/reference`` exists only in
the build container, not on the GPU box, so the seven source files of the path (the ones ``oracle/ref_shim.py`` executes
unmodified) are copied VERBATIM into ``oracle/_ref/`` -- git-ignored, so they never enter the history, but not
gpurun-ignored, so they travel to the GPU box next to the built ``.so`` files.  ``bench.py --impl reference`` and the
``cpu_baseline`` leg then time the reference's numba-parallel ``calc_gamma`` / ``calc_doppler_width`` /
``calc_alan_entries`` / ``single_theta_trace_parallel`` on the box's host cores (``oracle/ref_leg.py``).

    python -m oracle.stage_ref            # also run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("STARDIS_REFERENCE_SOURCE", "/root/reference")
BASE = "stardis/radiation_field/"
FILES = [
    BASE + "opacities/opacities_solvers/voigt.py",
    BASE + "opacities/opacities_solvers/broadening.py",
    BASE + "opacities/opacities_solvers/util.py",
    BASE + "opacities/opacities_solvers/base.py",
    BASE + "opacities/base.py",
    BASE + "source_functions/blackbody.py",
    BASE + "radiation_field_solvers/base.py",
]


def staged() -> bool:
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


def stage(verbose: bool = False) -> bool:
    """Copy the files when the reference tree is present; returns whether ``oracle/_ref`` is complete afterwards."""
    if os.path.isdir(os.path.join(SOURCE, "stardis", "radiation_field")):
        for f in FILES:
            dst = os.path.join(DEST, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(SOURCE, f), dst)
        if verbose:
            print(f"staged {len(FILES)} reference modules into {DEST}")
    return staged()


if __name__ == "__main__":
    print("complete" if stage(verbose=True) else "reference tree not found and oracle/_ref incomplete")
