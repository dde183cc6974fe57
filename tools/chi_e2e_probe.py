"""chi tail end to end with nu from pageable / page-locked host memory (development tool; the legs of bench.chi_leg)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
ctx = bench.Ctx()
out = bench.chi_leg(ctx)
print(json.dumps({k: out[k] for k in ("roofline", "e2e", "e2e_pinned_source") if k in out}, indent=1))
