"""Forward-only candidate evaluation on the resident sr_tiny super-network (BASELINE configs[4] shape: sampled sub-networks, forward
only, synthetic sub-val, bs 256).  Prints one JSON line: images/s over all candidates (device events), ms per candidate switch.
Usage: PYTHONPATH=. python tools/evo_eval_bench.py [--candidates 8] [--batches 6] [--space sr_tiny]"""
import argparse
import json
import random
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--space', default='sr_tiny')
    ap.add_argument('--candidates', type=int, default=8)
    ap.add_argument('--batches', type=int, default=6)
    ap.add_argument('--batch', type=int, default=256)
    args = ap.parse_args()
    from vit_search_b200 import supernet_config as sc, _lib
    from vit_search_b200.nets import create_model
    from vit_search_b200.evo_eval import CandidateEvaluator, sample_candidate
    nd, ks = sc.network_def(args.space), sc.num_channels_to_keep(args.space)
    torch.manual_seed(0)
    model = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=nd, num_classes=1000).cuda().eval()
    ev = CandidateEvaluator(model, 'cuda')
    g = torch.Generator(device='cuda').manual_seed(1)
    loader = [(torch.randn(args.batch, 3, 224, 224, device='cuda', generator=g), torch.randint(0, 1000, (args.batch,), device='cuda', generator=g))
              for _ in range(2)]                                     # two 154 MB batches alternate: inputs never sit in the 126 MB L2
    loader = [loader[i % 2] for i in range(args.batches)]
    rng = random.Random(0)
    cands = [sample_candidate(nd, ks, rng) for _ in range(args.candidates)]
    ev.score(cands[0], loader[:2])                                   # warm-up (operand copies, workspaces)
    ev.score(nd, loader[:2])
    torch.cuda.synchronize()
    lib = _lib.lib()
    l0 = lib.vsx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    scores = [ev.score(c, loader) for c in cands]
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    n_img = args.candidates * args.batches * args.batch
    t1 = time.perf_counter()
    for c in cands:
        model.set_active_subnet(c)
    model.set_active_subnet(None)
    switch_ms = (time.perf_counter() - t1) / len(cands) * 1e3
    # the largest network for scale
    e0.record()
    ev.score(nd, loader)
    e1.record()
    torch.cuda.synchronize()
    full_ms = e0.elapsed_time(e1)
    print(json.dumps({'metric': 'candidate evaluation images/sec (forward only, resident super-network weights)', 'space': args.space,
                      'value': n_img / ms * 1e3, 'unit': 'img/s', 'candidates': args.candidates, 'images_per_candidate': args.batches * args.batch,
                      'ms_per_batch': ms / (args.candidates * args.batches), 'wall_s': wall, 'host_ms_per_candidate_switch': switch_ms,
                      'largest_network_img_s': args.batches * args.batch / full_ms * 1e3, 'gpu_launches': lib.vsx_launch_count() - l0,
                      'acc1_first': scores[0]['acc1'], 'dtype': 'bf16', 'data': 'synthetic, device resident'}))


if __name__ == '__main__':
    main()
