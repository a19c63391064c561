import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import test_gpu_conv_tc as T
from helpers import norm_err
import helpers
# monkeypatch assertion to print
orig = T.norm_err
vals = []
def ne(a, b):
    v = orig(a, b); vals.append(v); return v
T.norm_err = ne
for args, kw in [((2,5,3,256,256,1),{}), ((2,6,3,256,256,1),dict(residual=True)), ((2,7,5,128,128,1),{}), ((3,9,10,64,64,1),{}), ((2,13,20,32,32,1),{}), ((8,16,3,256,256,1),{})]:
    vals.clear()
    try:
        T._case(*args, **kw)
        print(args, "ok", vals)
    except AssertionError as e:
        print(args, "FAIL", vals)
