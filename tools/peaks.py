"""FP64 ceilings of this GPU: register-resident DMMA (m8n8k4) and DFMA loops -> TFLOP/s.
MEASURED_PEAKS.json has no FP64 figure; bench.py uses the DMMA number as the tensor roofline."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psi4_b200 import Engine  # noqa: E402

e = Engine(1)
out = {"dmma_tflops": e.fp64_peak(0), "dfma_tflops": e.fp64_peak(1)}
print(json.dumps(out))
