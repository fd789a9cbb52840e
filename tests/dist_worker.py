"""Worker of tests/test_dist_nccl_gpu.py (also runnable by hand under torchrun): data-parallel TrainStep + GradSync on
NCCL vs the single-GPU full-batch gradient.

Every rank builds the same seeded model, takes rows [r*B_loc, (r+1)*B_loc) of ONE global batch with ONE global mask
set ("global-mask mode"), runs forward + backward + the gradient all-reduce exactly as
TrainStep.__call__ does, and compares its (averaged) flat gradients with the gradients rank 0's second model copy
gets from the full global batch on one GPU.  Prints one JSON line per configuration; exits non-zero on failure."""
import json
import os
import sys

import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from avjepa_b200.dist import GradSync, init_distributed
    from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator
    from helpers import MASK_CFG, build_product
    import step_support as S
    world, rank = init_distributed()
    dev = torch.device('cuda', torch.cuda.current_device())
    model = os.environ.get('AVJ_TEST_MODEL', 'vit_tiny')
    b_loc = 2
    B = b_loc * world
    g = torch.Generator().manual_seed(99)
    clips = torch.randn(B, 3, 16, 224, 224, generator=g)
    asgram = -80.0 * torch.rand(B, 1, 128, 192, generator=g)
    torch.manual_seed(234)
    coll = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    fake = [([torch.zeros(1)], 0, [0], torch.zeros(1)) for _ in range(B)]
    while True:
        try:
            _, ev, ea, pv, pa = coll(fake)
            break
        except TypeError:
            continue
    masks = dict(ev=ev, ea=ea, pv=pv, pa=pa)
    rows = slice(rank * b_loc, (rank + 1) * b_loc)
    loc = {k: [m[rows].contiguous().to(dev) for m in v] for k, v in masks.items()}
    full = {k: [m.to(dev) for m in v] for k, v in masks.items()}
    ok = True
    for mixed, tol in ((False, 1e-4), (True, 1e-2)):
        # single-GPU full-batch gradient (every rank computes it: cheap, and no broadcast needed)
        enc, pred = build_product(model, seed=0, device=dev, pred_depth=2)
        ref_step = S.make_train_step(enc, pred, mixed)
        loss_full, _, _ = ref_step.forward_loss(clips.to(dev), asgram.to(dev), full['ev'], full['ea'], full['pv'], full['pa'])
        loss_full.backward()
        ref = [gf.clone() for gf in ref_step.optimizer.flat_grads()]
        for overlap in ((False, True, 'bf16') if mixed else (False, True)):
            enc, pred = build_product(model, seed=0, device=dev, pred_depth=2)
            sync = GradSync(world, overlap=bool(overlap), layers_per_bucket=3)
            sync.bf16_wire = overlap == 'bf16'             # gradient intervals cross the wire as bf16 (AVJ_DDP_BF16=1)
            step = S.make_train_step(enc, pred, mixed, grad_sync=sync)
            loss, _, _ = step.forward_loss(clips[rows].to(dev), asgram[rows].to(dev), loc['ev'], loc['ea'], loc['pv'], loc['pa'])
            inv = step.backward_and_reduce(loss)
            got = [gf * inv for gf in step.optimizer.flat_grads()]
            num = sum(float(((a.double() - b.double()) ** 2).sum()) for a, b in zip(got, ref))
            den = sum(float((b.double() ** 2).sum()) for b in ref)
            err = (num / den) ** 0.5
            lsum = loss.detach().clone()
            tdist.all_reduce(lsum)
            lerr = abs(float(lsum) / world - float(loss_full)) / abs(float(loss_full))
            if rank == 0:
                print(json.dumps(dict(world=world, model=model, mixed=mixed, overlap=overlap, overlapped_active=bool(sync.overlap),
                                      grad_rel_err=err, loss_rel_err=lerr, tol=tol)), flush=True)
            ok = ok and err <= tol and lerr <= tol
            # and one full data-parallel step (optimizer + EMA) must keep the ranks bit-identical
            out = step(clips[rows].to(dev), asgram[rows].to(dev), loc['ev'], loc['ea'], loc['pv'], loc['pa'])
            w = dict(step.encoder.named_parameters())['backbone.blocks.0.attn.qkv.weight'].detach().clone()
            w0 = w.clone()
            tdist.broadcast(w0, 0)
            ok = ok and bool(torch.equal(w, w0)) and out[0] == out[0]
        # the pipelined finish (optimizer per reduced gradient interval, under the later all-reduces) must leave exactly the
        # state the all-at-once finish leaves: same kernels on the same values, only launched per interval
        states = []
        for pipe in (True, False):
            enc, pred = build_product(model, seed=0, device=dev, pred_depth=2)
            step = S.make_train_step(enc, pred, mixed, grad_sync=GradSync(world, overlap=True, layers_per_bucket=3))
            step.pipeline_optimizer = pipe
            for _ in range(2):
                step(clips[rows].to(dev), asgram[rows].to(dev), loc['ev'], loc['ea'], loc['pv'], loc['pa'])
            states.append([p.detach().clone() for m in (step.encoder, step.predictor, step.target_encoder) for p in m.parameters()])
        num = sum(float(((a.double() - b.double()) ** 2).sum()) for a, b in zip(*states))
        den = sum(float((b.double() ** 2).sum()) for b in states[1])
        perr = (num / den) ** 0.5
        same = all(torch.equal(a, b) for a, b in zip(*states))
        if rank == 0:
            print(json.dumps(dict(world=world, model=model, mixed=mixed, pipelined_optimizer_vs_plain_rel_err=perr, bit_identical=same)), flush=True)
        ok = ok and (same if not mixed else perr < 1e-4)
    tdist.barrier()
    tdist.destroy_process_group()
    if not ok:
        sys.exit(3)
    print('rank', rank, 'ok', flush=True)


if __name__ == '__main__':
    main()
