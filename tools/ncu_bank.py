"""One launch of the filterbank kernel per requested variant (for `ncu -k regex:bank_`):
python tools/ncu_bank.py C block nblk noct kernel pack logch [kernel pack logch ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from friture_b200.octavefilters import Octave_Filters  # noqa: E402

C, block, nblk, noct = (int(v) for v in sys.argv[1:5])
x = torch.randn((C, block * nblk), dtype=torch.float32, device="cuda") * 0.1
rest = sys.argv[5:]
for i in range(0, len(rest), 3):
    os.environ["FRT_BANK_KERNEL"] = rest[i]
    os.environ["FRT_BANK_PACK"] = rest[i + 1]
    os.environ["FRT_BANK_LOGCH"] = rest[i + 2]
    bank = Octave_Filters(3, n_octaves=noct)
    bank.energies_batch(x, block=block, db=True)
    torch.cuda.synchronize()
