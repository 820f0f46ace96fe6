"""One device-resident solve at half-dimension n (no warm-up, no checks): the process ncu wraps for a launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/ncu_one_solve.py n`).
Numbers printed by a run under ncu are not bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zquatev_b200 as z  # noqa: E402
from bench import make_input  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
left0 = make_input(n, dev)
work = torch.empty((2 * n, 2 * n), dtype=torch.complex128, device=dev)
work[:n].copy_(left0)
eig = torch.zeros(n, dtype=torch.float64, device=dev)
info = z.zquatev_device(2 * n, work.data_ptr(), 2 * n, eig.data_ptr())
print("info", info, z.last_phases())
