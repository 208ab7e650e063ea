"""Run-to-run reproducibility of the reference Actor's episode step on the new Executor (one GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import t2onet_b200 as T
from oracle import ref_shims
dev = torch.device('cuda', 0)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    opt = ref_shims.actor_options()
    actor = ref_shims.build_actor(opt, T.Executor, seed=10).to(dev)
    mod = bench._EpisodeL1(actor, opt)
    g = torch.Generator().manual_seed(100)
    B = 8
    x = torch.randint(4, 200, (B, opt.encoder_max_len), generator=g)
    x[:, 0], x[:, -1] = opt.start_id, opt.end_id
    img = torch.rand(B, 3, 32, 32, generator=g).to(dev)
    tgt = torch.rand(B, 3, 32, 32, generator=g).to(dev)
    x = x.to(dev)
    params = list(mod.parameters())
    outs = []
    for rep in range(3):
        mod.zero_grad(set_to_none=True)
        torch.manual_seed(7)
        loss = mod(x, img, tgt)
        loss.backward()
        used = [p for p in params if p.grad is not None]
        outs.append((loss.item(), torch.cat([p.grad.flatten() for p in used]).clone(), len(used)))
    for rep in (1, 2):
        d = (outs[rep][1] - outs[0][1]).abs().max() / outs[0][1].abs().max()
        print('tf32', tf32, 'rep', rep, 'loss', outs[rep][0], outs[0][0], 'grad rel diff', float(d), 'n used', outs[rep][2], outs[0][2])
    print('training mode:', actor.training, 'bn modules in train:', sum(m.training for m in actor.modules() if isinstance(m, torch.nn.BatchNorm2d)))
