"""which call of the backward pass invalidates CUDA-graph capture?  (prints the capture status after every C-ABI call)"""
import ctypes
import sys
import threading
import torch
sys.path.insert(0, '/root/repo')
from tests import util
from audiossl_b200 import _lib, ops
from audiossl_b200.methods.atst.model import ATSTLightningModule

cudart = ctypes.CDLL("libcudart.so.12")
state = {"on": False, "last": None, "reported": False, "n": 0}
orig_check = _lib.check


def status():
    st = ctypes.c_int(0)
    rc = cudart.cudaStreamIsCapturing(ctypes.c_void_p(_lib.stream()), ctypes.byref(st))
    return rc, st.value


def check(rc, what=""):
    orig_check(rc, what)
    if state["on"] and not state["reported"]:
        state["n"] += 1
        rc2, st = status()
        if st != 1 or rc2 != 0:
            state["reported"] = True
            print("capture status %d (rc %d) after call #%d %s on thread %s; previous ok call: %s" % (
                st, rc2, state["n"], what, threading.current_thread().name, state["last"]))
        else:
            state["last"] = "%s [%s]" % (what, threading.current_thread().name)


_lib.check = check
ops.check = check

torch.manual_seed(0)
lm = ATSTLightningModule(arch=dict(embed_dim=128, depth=2, num_heads=2), learning_rate=1e-3, warmup_steps=2, max_steps=20,
                         ema=0.9, drop_path_rate=0.0)
util.load_det(lm.model)
lm.cuda().train()
opt = lm.configure_optimizers()[0]
lm.trainer.optimizers = [opt]
crops, lengths = util.make_inputs("graph0", 8, [101, 101], [[101 - (i * 3) % 30 for i in range(8)], [101] * 8])
batch = (([c.cuda() for c in crops], [l.cuda() for l in lengths]), None)


def fwd_bwd():
    loss = lm.training_step(batch, 0)
    opt.zero_grad()
    print("   status before backward:", status() if state["on"] else "-")
    loss.backward()
    print("   status after backward:", status() if state["on"] else "-")
    return loss.detach()


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        fwd_bwd()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        state["on"] = True
        out = fwd_bwd()
        state["on"] = False
    print("captured ok")
except Exception as e:  # noqa: BLE001
    import traceback
    traceback.print_exc()
