import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from test_bank_gpu import make_x, oracle_run
from friture_b200.octavefilters import Octave_Filters
np.set_printoptions(linewidth=250, precision=3)
for (C, T, block, pack, logch) in ((1, 1024, 256, 1, 5), (2, 1024, 256, 1, 5), (3, 2048, 512, 1, 6), (2, 1024, 256, 2, 5)):
    os.environ["FRT_BANK_KERNEL"]="pipe"; os.environ["FRT_BANK_PACK"]=str(pack); os.environ["FRT_BANK_LOGCH"]=str(logch)
    x = make_x(C, T, seed=1)
    bank = Octave_Filters(3)
    e = bank.energies_batch(torch.from_numpy(x).cuda(), block=block).cpu().numpy().astype(np.float64)
    E, _ = oracle_run(bank, x, block)
    print("C",C,"T",T,"block",block,"pack",pack,"logch",logch)
    for c in range(C):
        for b in range(min(4, T//block)):
            print(" ch",c,"blk",b,"ratio got/ref:", e[c,b]/E[c,b])
