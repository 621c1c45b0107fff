"""Development sweep: thread-per-sequence kernels under different CTA sizes / occupancy pads (env knobs are read once per
process, so every setting is its own subprocess).  python tools/time_seq_knobs.py B T"""
import os, subprocess, sys
B = sys.argv[1] if len(sys.argv) > 1 else "262144"
T = sys.argv[2] if len(sys.argv) > 2 else "20"
for fw, bw, pad in [(4, 2, 0), (4, 2, 20000), (2, 2, 0), (2, 2, 12000), (4, 2, 50000)]:
    env = dict(os.environ, KVAE_SEQ_FWD_WARPS=str(fw), KVAE_SEQ_BWD_WARPS=str(bw), KVAE_SEQ_SMEM_PAD=str(pad))
    r = subprocess.run([sys.executable, "tools/time_large.py", B, T, "1"], env=env, capture_output=True, text=True)
    out = r.stdout
    if r.returncode != 0:
        print(f"fwd_warps={fw} bwd_warps={bw} pad={pad}: FAILED rc={r.returncode}: {r.stderr[-400:]}", flush=True)
    for line in out.splitlines():
        if "L=1" in line:
            print(f"fwd_warps={fw} bwd_warps={bw} pad={pad}: {line}", flush=True)
