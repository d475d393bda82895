#!/usr/bin/env python
"""Build a tuning variant of the library next to the product one:  python tools/build_variant.py NAME -DFLAG=1 ...
-> audiodeepfake-detection_b200/libafd_b200_NAME.so (for tools/ab_bench.py / tools/wpt_phase_timing.py)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_afd_build", os.path.join(ROOT, "audiodeepfake-detection_b200", "build.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
only = [a[7:] for a in sys.argv[2:] if a.startswith("--only=")]
flags = [a for a in sys.argv[2:] if not a.startswith("--only=")]
if only:
    # recompile just these units with the flags; every other object is taken from the product build
    mod.SOURCES_VARIANT_ONLY = only[0].split(",")
print(mod.build(True, verbose="-v" in flags, extra_flags=[f for f in flags if f != "-v"],
                out_path=os.path.join(ROOT, "audiodeepfake-detection_b200", f"libafd_b200_{sys.argv[1]}.so")))
