"""Run a short training with K sharded over WORLD_SIZE ranks; rank 0 prints the loss_log as JSON.
Used by tests/test_multi_gpu.py: the result must not depend on the number of ranks (global-index Philox)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt
import torch.distributed as td
import pspde

world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
pt.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=pt.device("cuda", local))
loss = sys.argv[1] if len(sys.argv) > 1 else "log-variance"
detach = loss != "relative_entropy"
d = 10
prob = pspde.LLGC(d=d, T=0.5, device="cuda")
S = pspde.Solver("ddp", prob, K=1000, L=4, lr=1e-2, delta_t=0.05, time_approx="inner", loss_method=loss,
                 detach_forward=detach, u_l2_error_flag=False, early_stopping_time=None, verbose=False, seed=5)
S.z_n = pspde.DenseNet(d_in=d + 1, d_out=d, lr=1e-2, seed=42)
S.update_Phis()
S.train()
if int(os.environ.get("RANK", "0")) == 0:
    print("DDP_RESULT " + json.dumps({"world": world, "loss_log": S.loss_log, "theta_sum": float(S._theta.sum()),
                                      "K_local": S._get_engine().K_local}))
if world > 1:
    td.destroy_process_group()
