"""FP64 ceilings of this GPU: register-resident DMMA (m8n8k4) and DFMA loops -> TFLOP/s.
MEASURED_PEAKS.json has no FP64 figure; bench.py uses the DMMA number as the tensor roofline."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psi4_b200 import Engine  # noqa: E402

e = Engine(1)
out = {"dmma": e.fp64_peak(0, 2.0), "dfma": e.fp64_peak(1, 2.0), "dmma+dfma": e.fp64_peak(2, 2.0),
       "dmma_8warps_per_sm": e.fp64_peak(3, 1.0), "dmma_4warps_per_sm": e.fp64_peak(4, 1.0)}
print(json.dumps(out))
