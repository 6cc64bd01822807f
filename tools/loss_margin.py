import numpy as np, torch, sys
sys.path.insert(0, ".")
import derl_b200 as d
from oracle import derl_oracle as O
K = torch.ops.derl_b200
def cuda(x): return torch.from_numpy(np.ascontiguousarray(x)).cuda()
for kind, width in (("categorical",4),("categorical",18),("categorical",130),("gaussian",6),("gaussian",40)):
  nb = 131072 if width <= 18 else 20000
  rng = np.random.RandomState(width)
  batch = dict(log_prob=(rng.standard_normal(nb) * .2 - 1.5).astype(np.float32),
               advantages=rng.standard_normal(nb).astype(np.float32),
               value_targets=rng.standard_normal((nb, 1)).astype(np.float32),
               values=rng.standard_normal((nb, 1)).astype(np.float32))
  pred = (batch["values"] + rng.standard_normal((nb, 1)) * .15).astype(np.float32)
  if kind == "categorical":
    head = [rng.standard_normal((nb, width)).astype(np.float32)]
    batch["actions"] = rng.randint(0, width, nb).astype(np.int64)
  else:
    head = [rng.standard_normal((nb, width)).astype(np.float32), np.exp(rng.standard_normal((nb, width)) * .2).astype(np.float32)]
    batch["actions"] = (head[0] + .5 * rng.standard_normal((nb, width))).astype(np.float32)
    batch["log_prob"] = (batch["log_prob"] - width).astype(np.float32)
  ch = [torch.tensor(h, requires_grad=True) for h in head]; cv = torch.tensor(pred, requires_grad=True)
  want = O.ppo_loss(ch, cv, batch, 0.1, 0.25, 0.01); want.backward()
  # fp64 truth
  lw, gw = O.ppo_loss_closed_form(ch, cv, batch, 0.1, 0.25, 0.01)
  dh = [cuda(h).requires_grad_() for h in head]; dv = cuda(pred).requires_grad_()
  if kind == "categorical":
    loss, dl, dvv, st = K.ppo_loss_categorical(dh[0], dv, cuda(batch["actions"]), cuda(batch["log_prob"]), cuda(batch["advantages"]), cuda(batch["value_targets"]), cuda(batch["values"]), 0.1, 0.25, 0.01)
    grads = [dl, dvv]
  else:
    loss, dl, ds, dvv, st = K.ppo_loss_gaussian(dh[0], dh[1], dv, cuda(batch["actions"]), cuda(batch["log_prob"]), cuda(batch["advantages"]), cuda(batch["value_targets"]), cuda(batch["values"]), 0.1, 0.25, 0.01)
    grads = [dl, ds, dvv]
  print(kind, width, "loss gpu", loss.item(), "cpu32", want.item(), "fp64", lw, "rel(gpu,cpu32)", abs(loss.item()-want.item())/abs(want.item()), "rel(cpu32,fp64)", abs(want.item()-lw)/abs(lw))
  for g, r in zip(grads, [h.grad for h in ch] + [cv.grad]):
    r = r.numpy(); e = np.abs(g.detach().cpu().numpy() - r); tol = 1e-5*np.abs(r) + 1e-5*np.abs(r).max()
    print("   grad worst err/tol", float((e/tol).max()))
