"""cuDNN GRU vs ATen's native GRU (cudnn disabled) at the C2 shape: forward + backward time, and agreement."""
import sys, time, torch
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 205_587
T, I, H = 16, 6, 96
torch.manual_seed(0)
gru = torch.nn.GRU(I, H, 1, batch_first=True).to(dev)
x = torch.randn(N, T, I, device=dev)
up = torch.randn(N, H, device=dev)

def run(enabled, reps=3):
    with torch.backends.cudnn.flags(enabled=enabled):
        def step():
            gru.zero_grad()
            h = gru(x)[1][-1]
            h.backward(up)
            return h.detach(), [p.grad.clone() for p in gru.parameters()]
        step(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = step()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, out

for en in (True, False):
    ms, (h, grads) = run(en)
    print("cudnn" if en else "native", f"{ms:.2f} ms fwd+bwd", f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    if en:
        ref = (h, grads)
    else:
        print("max |dh|", float((h - ref[0]).abs().max()), "max |dgrad| rel", max(float((g - r).abs().max() / r.abs().max()) for g, r in zip(grads, ref[1])))
