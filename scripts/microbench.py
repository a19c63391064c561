"""Per-kernel device times (CUDA events, warm) at the bench shapes.  Usage: python scripts/microbench.py [B]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aesrc2020_b200 import ops, tc, model as mdl, utils as us
import io, contextlib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 500
dev = torch.device("cuda")
with contextlib.redirect_stdout(io.StringIO()):
    model, _ = mdl.SAR_Net((T, 80, 1), ctc_enable=True, disc_enable=True, res_type="res34", res_filters=32, mto="gvlad",
                           vlad_clusters=64, ghost_clusters=8, metric_loss="arcface")
eng = model.engine()
x, _ = us.synthetic_batch(model.config, B, seed=1)
xd = {k: model._to_device(k, v).clone() for k, v in x.items()}


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


res = {}
p = eng.p
S = eng.plan.seq_len
rn = eng.resnet
res["forward_total"] = timeit(lambda: eng.forward(xd))
res["resnet_total"] = timeit(lambda: rn.forward(xd["x_data"]))
cur = rn._buf(0, B, eng.plan.pool_hout, eng.plan.pool_wout, 64, False, "raw")
s_, t_ = rn.bn(eng.plan.stem.post_bn)
res["stem_pool"] = timeit(lambda: tc.stem_pool(xd["x_data"], rn.p["stem/kernel"], rn.p["stem/bias"], s_, t_, cur))
seq = torch.randn(B, S, 256, device=dev)
seq512 = torch.randn(B, S, 512, device=dev)
res["dense_256_256_tanh"] = timeit(lambda: ops.dense(seq, p["CNN_LIN/kernel"], p["CNN_LIN/bias"], act="tanh"))
res["dense_512_256_tanh"] = timeit(lambda: ops.dense(seq512, p["AR_DS/kernel"], p["AR_DS/bias"], act="tanh"))
res["layernorm_256"] = timeit(lambda: ops.layernorm(seq, p["CNN_LIN_LN/gamma"], p["CNN_LIN_LN/beta"]))
res["gru_proj_256_1536"] = timeit(lambda: ops.dense(seq, p["CRNN/kernel_cat"], p["CRNN/ibias_cat"]))
res["gru_proj_512_1536"] = timeit(lambda: ops.dense(seq512, p["CTC_BIGRU/kernel_cat"], p["CTC_BIGRU/ibias_cat"]))
xp = torch.randn(B, S, 2, 768, device=dev) * 0.1
res["bigru_recurrence"] = timeit(lambda: ops.bigru(xp, p["CRNN/rec"], p["CRNN/rbias"], seq=True))
res["vlad"] = timeit(lambda: ops.vlad(seq, p["gvlad/w_assign"], p["gvlad/b_assign"], p["gvlad/centers"], 64, 8))
xpl = tc.Planes((torch.randn(2, B * S, 256, device=dev) * 0.5).half(), 1, B * S, 1, 256, False)
vpl = tc.alloc_rows(B, 64 * 256, dev)
res["vlad_tc"] = timeit(lambda: tc.vlad_tc(xpl, p["gvlad/w_assign_tc"], p["gvlad/b_assign"], p["gvlad/centers"], B, S, 64, 8, planes=vpl, want_dense=False))
integ = torch.randn(B, 16384, device=dev)
res["embed_splitk"] = timeit(lambda: eng.embed(integ))
emb = torch.randn(B, 256, device=dev)
cls = tuple(p[k] for k in ("AR_CF_DS1/kernel", "AR_CF_DS1/bias", "AR_CF_DS2/kernel", "AR_CF_DS2/bias", "y_accent/kernel", "y_accent/bias"))
res["head"] = timeit(lambda: ops.head(emb, cls, wd=p["y_disc/w"], onehot=xd["x_accent"], n_classes=8, head_kind="arcface"))
logits = torch.randn(B, S, 1000, device=dev)
res["ctc_pred_gemm"] = timeit(lambda: ops.dense(seq, p["ctc_pred/kernel"], p["ctc_pred/bias"]))
res["ctc"] = timeit(lambda: ops.ctc(logits, xd["x_ctc_label"], xd["x_ctc_in_len"], xd["x_ctc_out_len"]))
# per conv layer
for i, b in enumerate(eng.plan.blocks):
    pass
print(json.dumps({"B": B, "us": {k: round(v, 1) for k, v in res.items()}}))
