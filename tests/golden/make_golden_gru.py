"""Golden vectors for the sequence head, produced by the reference's own layer: `torch.nn.GRU` on the CPU, built and
called exactly as src/models/grusage.py:55-60 / :160-161 do (`gru_out, hlast = self.gru(x); x = hlast[-1]`).

    python tests/golden/make_golden_gru.py            # writes tests/golden/gru/*.pt

Each fixture: the constructor arguments, the state dict, x [N,T,I], an upstream gradient for the last hidden state, and
-- in fp32 and, for adjudication, fp64 -- the last hidden state, the gradient of every parameter and of x."""
import os

import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gru")
CASES = {"ref_c2_hidden96": dict(N=150, T=16, I=6, H=96),      # the reference's configuration (main.py:42-44)
         "ref_hidden64_wide_input": dict(N=70, T=9, I=8, H=64),
         "ref_hidden32_single_step": dict(N=33, T=1, I=3, H=32)}


def main():
    os.makedirs(OUT, exist_ok=True)
    for seed, (name, c) in enumerate(CASES.items()):
        torch.manual_seed(100 + seed)
        gru = torch.nn.GRU(input_size=c["I"], hidden_size=c["H"], num_layers=1, batch_first=True)
        x = torch.randn(c["N"], c["T"], c["I"])
        up = torch.randn(c["N"], c["H"])
        fix = dict(cfg=c, state_dict={k: v.clone() for k, v in gru.state_dict().items()}, x=x, up=up)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            g = torch.nn.GRU(input_size=c["I"], hidden_size=c["H"], num_layers=1, batch_first=True).to(dt)
            g.load_state_dict({k: v.to(dt) for k, v in gru.state_dict().items()})
            xi = x.detach().clone().to(dt).requires_grad_(True)
            _, hlast = g(xi)
            h = hlast[-1, :, :]
            h.backward(up.to(dt))
            fix[tag] = dict(h=h.detach().clone(), dx=xi.grad.clone(), grads={k: p.grad.clone() for k, p in g.named_parameters()})
        torch.save(fix, os.path.join(OUT, name + ".pt"))
        print(name, {k: tuple(v.shape) for k, v in fix["f32"]["grads"].items()})


if __name__ == "__main__":
    main()
