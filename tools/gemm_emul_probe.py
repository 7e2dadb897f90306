import os, sys, torch, time
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(32, 2048, 128, device=dev)
w = torch.randn(128, 512, device=dev)
ref = (x.double() @ w.double())
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
f = lambda: torch.bmm(x, w.unsqueeze(0).expand(32, -1, -1))
out = f()
print(os.environ.get("CUBLAS_EMULATE_SINGLE_PRECISION"), "us", round(t(f), 1), "max rel err", float((out.double() - ref).abs().max() / ref.abs().max()))
torch.backends.cuda.matmul.allow_tf32 = True
out = f()
print("tf32", "us", round(t(f), 1), "max rel err", float((out.double() - ref).abs().max() / ref.abs().max()))
