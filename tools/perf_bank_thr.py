import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.getcwd())
from tools.perf_bank import run
for C, nb in ((2048, 128), (4096, 64), (8192, 64), (16384, 32)):
    run(C, 512, nb)
