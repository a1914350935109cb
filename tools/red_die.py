"""python tools/red_die.py <mode>: one launch of the reduction-placement hook (run under ncu, tools/deposit_variants.sh)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcxcl_b200 import abi

lib = abi.load()
n, tot = C.c_uint64(0), C.c_double(0)
abi.check(lib.mcxb_bench_red_die(0, int(sys.argv[1]), 148 * 4, 2000, C.byref(n), C.byref(tot)))
print("mode", sys.argv[1], "issued", n.value, "sum", tot.value)
