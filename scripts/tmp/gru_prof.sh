set -e
cd $GRAFT_REPO_ROOT
cp aesrc2020_b200/csrc/libsarnet_sm100.so /tmp/lib_keep.so
SAR_NVCC_EXTRA=-DSAR_GRU_PROFILE python -m aesrc2020_b200.csrc.build --force > /dev/null 2>&1
python - <<'P'
import torch, sys
sys.path.insert(0, '.')
from aesrc2020_b200 import _shim
from aesrc2020_b200._shim import ptr, stream_ptr
B,S,U=64,48,256
xp = torch.randn(B, S, 2, 3 * U, device="cuda") * 0.1
rec = torch.randn(2, U, 3 * U, device="cuda") * 0.05
rb = torch.randn(2, 3 * U, device="cuda") * 0.1
out = torch.empty(B, S, 2 * U, device="cuda")
lib=_shim.lib()
for flags in (1,1,7):
    print("flags", flags, flush=True)
    lib.sar_bigru_fwd(ptr(xp), ptr(rec), ptr(rb), ptr(out), B, S, U, flags, stream_ptr())
    torch.cuda.synchronize()
P
cp /tmp/lib_keep.so aesrc2020_b200/csrc/libsarnet_sm100.so
