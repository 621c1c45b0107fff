"""lstm dynamics + missing observations (imputation with the original KVAE): the fused launch (LSTM cell inside the
filter kernel, kvae_kf_filter_lstm_fwd) against the per-step path (cuDNN LSTM step + one filter launch per time step)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import KalmanFilter, DynamicsParameter

dev = torch.device("cuda:0")
n, p, m, K = 4, 2, 4, 3
torch.manual_seed(0)
A = torch.eye(n).repeat(K, 1, 1) + 0.05 * torch.randn(K, n, n)
dyn = DynamicsParameter(A, 0.05 * torch.randn(K, n, m), 0.3 * torch.randn(K, p, n), hidden_lstm=50)
kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, torch.zeros(n), 20.0 * torch.eye(n), dyn).to(dev).eval()
fused_impl = kf._run_fused_lstm


def run(B, T, fused, reps):
    Y = torch.randn(B, T, p, device=dev)
    U = torch.zeros(B, T, m, device=dev)
    mask = (torch.rand(B, T, device=dev) < 0.5).float()
    kf._run_fused_lstm = fused_impl if fused else (lambda *a, **k: None)
    with torch.no_grad():
        for _ in range(2):
            dyn.reset_state(); kf.smooth(Y, U, mask)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dyn.reset_state(); kf.smooth(Y, U, mask)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for B, T in ((8192, 20), (8192, 200), (65536, 200)):
    tf = run(B, T, True, 5)
    ts = run(B, T, False, 2 if T >= 200 else 5)
    print(f"B={B} T={T}: fused {tf * 1e3:.2f} ms ({B * T / tf / 1e9:.2f} G seq-steps/s), per-step path {ts * 1e3:.1f} ms "
          f"({B * T / ts / 1e9:.3f} G seq-steps/s): {ts / tf:.0f}x", flush=True)
