"""Test infrastructure: bench.py's b200 arm with the CPU stand-in swapped in for the product library and a pretended device (tests/test_bench_script.py).
Checks the script's own code paths; the numbers mean nothing and are never reported."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arriba_b200 import lib as L, _build
hs=_build.build_hostsim()
torch.cuda.is_available=lambda: True
torch.cuda.synchronize=lambda *a,**k: None
_build.PRODUCT_LIB=hs
orig_load=L.load
L.load=lambda path=None: orig_load(hs)
OrigP=L.Pipeline
class P(OrigP):
    def __init__(self,*a,**k):
        k.setdefault('lib_path',hs); super().__init__(*a,**k)
L.Pipeline=P
import bench
sys.argv=['bench.py']+sys.argv[1:]
bench.main()
