"""CPU: host-side logic of the product (no kernels run): sampling protocol, module surface, registry, MAC counter,
C-ABI export table, and the loud failure when asked to compute without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import vit_res_oracle as O
from oracle.cases import CASES, SMALL_DEF, SMALL_SPACE, VIT_RES_TINY
from vit_search_b200 import _lib, core, macs
from vit_search_b200 import supernet_config as sc
from vit_search_b200.nets import (Block, ChannelDrop, FlexibleDistillVisionTransformerSR, MaskedLayerNorm, create_model,
                                  list_models)

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small(**kw):
    return create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=SMALL_DEF, num_classes=1000, drop_rate=0.,
                        drop_path_rate=0., drop_block_rate=None, num_channels_to_keep=SMALL_SPACE, **kw)


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'vsx.h')).read()
    declared = set(re.findall(r'\b(vsx_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'vsx_gemm_desc', 'vsx_adamw_tensor'}
    lib = _lib.lib()                      # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), 'libvsx.so does not export %s' % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert lib.vsx_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback():
    blk = Block(64, 2, 32, 128)
    with pytest.raises(RuntimeError, match='no CPU'):
        blk(torch.randn(2, 5, 64))
    with pytest.raises(RuntimeError, match='no CPU'):
        MaskedLayerNorm(64)(torch.randn(2, 5, 64))
    m = _small(example_per_arch=2, num_warmup_epochs=0)
    m.set_epoch(0)
    with pytest.raises(RuntimeError, match='no CPU'):
        m(torch.randn(8, 3, 224, 224))


def test_state_dict_surface_matches_reference_layout():
    for nd, sup in ((SMALL_DEF, True), (VIT_RES_TINY, False), (sc.network_def('sr_tiny_mh'), True)):
        if sup:
            space = SMALL_SPACE if nd is SMALL_DEF else sc.num_channels_to_keep('sr_tiny_mh')
            m = create_model('flexible_vit_sr_patch14_224_patch_output_supernet', network_def=nd, num_channels_to_keep=space,
                             example_per_arch=2, num_warmup_epochs=0)
        else:
            m = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=nd)
        sd, sh = m.state_dict(), O.param_shapes(nd)          # O.param_shapes is asserted against the reference in make_golden
        assert list(sd) == list(sh)
        assert all(tuple(sd[k].shape) == tuple(sh[k]) for k in sd)
    assert len(sd) == 258                                    # SURVEY.md §8b probe for sr_tiny_mh
    assert m.no_weight_decay() == {'tokens'}
    assert len([n for n in list_models() if n.startswith('flexible_vit_sr_')]) == 9


@pytest.mark.parametrize('name', [n for n, c in CASES.items() if c['supernet'] and c.get('train', True)])
def test_sampling_protocol_matches_reference(name):
    """Same CPU-RNG seed => same sub-architectures as the reference (keeps recorded in the golden files)."""
    case = CASES[name]
    m = _small(example_per_arch=case['epa'], num_warmup_epochs=case['warmup'], single_arch=case.get('single', False),
               hybrid_arch=case.get('hybrid', False))
    m.set_epoch(case['epoch'])
    m.train()
    torch.manual_seed(case['seed'])
    keeps = m.sample_keeps(case['batch'])
    flat = [k[n] for k in keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
    assert flat == np.load(os.path.join(GOLD, name + '.npz'))['keeps'].tolist()
    perm = m._group_permutation(keeps, case['batch'])
    if perm is not None:
        assert sorted(perm) == list(range(case['batch']))
        sig = [tuple(v[b] for k in keeps for v in k.values()) for b in perm]
        seen, last = set(), None
        for s in sig:                                        # groups are contiguous after the permutation
            assert s == last or s not in seen
            seen.add(s)
            last = s
    m.eval()
    full = m.sample_keeps(3)
    assert full[0]['embed'] == [SMALL_DEF[0][1]] * 3         # eval: all-true masks (nets/channel_drop.py:84-88)


def test_channel_drop_tables_and_epoch_reset():
    G = np.load(os.path.join(GOLD, 'functions.npz'))
    i = 0
    for epoch in (0, 2, 5, 9):
        for single in (False, True):
            cd = ChannelDrop(np.array([96, 64, 128, 32, 80]), num_warmup_epochs=5, example_per_arch=2, single_arch=single)
            cd.set_epoch(epoch)
            cd.train()
            torch.manual_seed(100 + epoch)
            assert cd.draw(12, 128) == G['cd_draws'][i].tolist()
            assert cd.keep_table == O.keep_table([96, 64, 128, 32, 80], 12, 2, single, epoch, 5)
            cd.set_epoch(epoch + 1)
            assert cd.keep_table is None
            i += 1
    cd = ChannelDrop(np.array([64, 32]), num_warmup_epochs=0, example_per_arch=4)
    cd.set_epoch(0)
    with pytest.raises(AssertionError):
        cd.draw(6, 64)                                       # batch not divisible by example_per_arch (reference :123)


def test_segments_and_keep_algebra():
    segs = core.make_segments(6, 64, [64, 64, 44, 44, 44, 64], [32, 32, 32, 64, 64, 64], 64, [64, 64, 0, 0, 44, 64])
    assert [(s.b0, s.b1, s.ek, s.ik, s.ck, s.active) for s in segs] == [
        (0, 2, 64, 32, 64, True), (2, 3, 44, 32, 0, False), (3, 4, 44, 64, 0, False), (4, 5, 44, 64, 44, True), (5, 6, 64, 64, 64, True)]
    from vit_search_b200.nets._masks import and_keep
    assert and_keep([4, 8], None) == [4, 8] and and_keep([4, 8], [6, 2]) == [4, 2] and and_keep(None, None) is None


def test_rewiring_sorts_by_l1_like_reference():
    torch.manual_seed(0)
    blk = Block(32, 4, 8, 64)
    sd0 = {k: v.clone() for k, v in blk.state_dict().items()}
    blk.rewiring()
    w1, b1, w2 = sd0['mlp.fc1.weight'], sd0['mlp.fc1.bias'], sd0['mlp.fc2.weight']
    score = w2.abs().sum(0) + w1.abs().sum(1) + b1.abs()
    idx = torch.sort(score, descending=True)[1]
    assert torch.equal(blk.mlp.fc1.weight.data, w1[idx]) and torch.equal(blk.mlp.fc2.weight.data, w2[:, idx])
    q = sd0['attn.qkv.weight']
    hs = q.abs().sum(1).reshape(3, 4, 8).sum((0, 2)) + sd0['attn.qkv.bias'].abs().reshape(3, 4, 8).sum((0, 2)) + \
        sd0['attn.proj.weight'].abs().sum(0).reshape(4, 8).sum(1)
    hidx = torch.sort(hs, descending=True)[1]
    assert torch.equal(blk.attn.qkv.weight.data, q.reshape(3, 4, 8, -1)[:, hidx].reshape(96, -1))
    assert torch.equal(blk.attn.proj.weight.data, sd0['attn.proj.weight'].reshape(-1, 4, 8)[:, hidx].reshape(-1, 32))


def test_mac_counter_known_answers():
    G = np.load(os.path.join(GOLD, 'functions.npz'))
    assert macs.network_macs(VIT_RES_TINY) == int(G['mac_vit_res_tiny']) == 1794378240     # compute_flop_mac.py __main__, tiny.sh:19
    assert macs.network_macs(SMALL_DEF) == int(G['mac_small_def'])
    assert macs.network_macs(sc.network_def('sr_tiny_mh')) == 3497553920
    assert macs.network_macs(sc.network_def('sr_tiny')) == 3650185728
    m = _small(example_per_arch=2, num_warmup_epochs=0)
    m.set_epoch(0)
    m.train()
    torch.manual_seed(3)
    keeps = m.sample_keeps(8)
    eff = [macs.network_macs(macs.effective_network_def(SMALL_DEF, keeps, b)) for b in range(8)]
    assert max(eff) <= macs.network_macs(SMALL_DEF) and min(eff) > 0


def test_fused_adamw_grouping_matches_timm_rules():
    from vit_search_b200.engine import FusedAdamW
    m = _small(example_per_arch=2, num_warmup_epochs=0)
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    wd = {n: w for n, _, w in opt.entries}
    assert wd['tokens'] == 0.0 and wd['pos_embed'] == 0.05 and wd['blocks.0.attn.qkv.weight'] == 0.05
    assert wd['blocks.0.attn.qkv.bias'] == 0.0 and wd['blocks.0.norm1.weight'] == 0.0 and wd['patch_embed.conv1.bn.weight'] == 0.0
    assert wd['blocks.2.pos_embed'] == 0.05


def test_subnet_extents_of_search_candidates():
    """Candidate definitions (search_utils/gen_utils.py form) -> prefix extents on the resident super-network; misfits are rejected."""
    import pytest
    from oracle.cases import EVO_SUPER_DEF, EVO_CANDIDATES
    from vit_search_b200.nets import create_model
    m = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=EVO_SUPER_DEF, num_classes=1000)
    e = m.subnet_extents(EVO_CANDIDATES['narrow'])
    assert e[0] == {'embed': 56} and e[1] == {'attn': 64, 'mlp': 96} and e[2] == {'attn': 32, 'mlp': 64}
    assert e[3] == {'embed': 112} and e[6] == {'skip': True} and e[7] == {'embed': 224} and e[-1] == {}
    full = m.subnet_extents(EVO_SUPER_DEF)
    assert full[8] == {'attn': 256, 'mlp': 512}
    bad = list(EVO_CANDIDATES['narrow'])
    bad[3] = (3, 64, 112)                                   # SR input width must follow the embedding width
    with pytest.raises(ValueError):
        m.subnet_extents(tuple(bad))
    with pytest.raises(ValueError):
        m.subnet_extents(EVO_CANDIDATES['narrow'][:-1])
    bad = list(EVO_CANDIDATES['narrow'])
    bad[8] = (1, (224, 3, 32), (224, 384), 1)               # head_dim differs from the super-network's 64
    with pytest.raises(ValueError):
        m.subnet_extents(tuple(bad))


def test_vit16_module_surface():
    """nets/vision_transformer_supernet.py drop-in: factories, state_dict keys / shapes (= the reference's, checked against it in
    oracle/make_golden_vit16.py), no_weight_decay, draw order."""
    import numpy as np
    import os
    import torch
    from oracle import vit_res_oracle as O
    from oracle.cases import VIT16_CASES, VIT16_DEF, VIT16_SPACE
    from vit_search_b200.nets import create_model, list_models
    for n in ('flexible_vit_patch16_224', 'flexible_vit_patch16_224_supernet', 'flexible_vit_patch16_192', 'flexible_vit_patch16_192_supernet'):
        assert n in list_models()
    case = VIT16_CASES['vit16_multi']
    m = create_model('flexible_vit_patch16_224_supernet', network_def=VIT16_DEF, num_classes=1000, num_channels_to_keep=VIT16_SPACE,
                     example_per_arch=case['epa'], num_warmup_epochs=case['warmup'])
    shapes = O.param_shapes(VIT16_DEF, num_tokens=2, patch_output=False, patch_size=16)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    assert m.no_weight_decay() == {'pos_embed', 'tokens'} and m.num_tokens == 2
    m.set_epoch(case['epoch'])
    m.train()
    torch.manual_seed(case['seed'])
    keeps = m.sample_keeps(case['batch'])
    flat = [k[n] for k in keeps for n in ('embed', 'attn', 'layer', 'mlp') if n in k]
    G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'vit16_multi.npz'))
    assert flat == G['keeps'].tolist()
    m192 = create_model('flexible_vit_patch16_192', network_def=VIT16_DEF, num_classes=1000)
    assert m192.pos_embed.shape == (1, 144 + 2, 64)


def test_abi_signatures_match_the_header():
    """Every ctypes signature in _lib.SIGNATURES has the arity and the scalar / pointer kinds of its prototype in include/vsx.h, and the
    ctypes mirrors of the descriptor structs have the header's field order (ABI drift would otherwise only show up as garbage on a GPU)."""
    import ctypes as C
    hdr = open(os.path.join(ROOT, 'include', 'vsx.h')).read()
    hdr = re.sub(r'/\*.*?\*/', ' ', hdr, flags=re.S)
    kinds = {C.c_void_p: 'p', C.c_int: 'i', C.c_long: 'l', C.c_float: 'f', C.c_double: 'd'}
    protos = dict(re.findall(r'\b(?:int|long|const char\*)\s+(vsx_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', hdr, flags=re.S))
    assert set(protos) >= set(_lib.SIGNATURES)
    for name, sig in _lib.SIGNATURES.items():
        params = [p.strip() for p in protos[name].replace('\n', ' ').split(',')]
        if params == ['void'] or params == ['']:
            params = []
        want = ''
        for p in params:
            if '*' in p:
                want += 'p'
            else:
                ty = p.split()[0] if not p.startswith('const ') else p.split()[1]
                want += {'int': 'i', 'long': 'l', 'float': 'f', 'double': 'd'}[ty]
        got = ''.join(kinds.get(t, 'p') for t in sig)          # POINTER(struct) arguments are pointers
        assert got == want, (name, got, want)
    for cname, cls in (('vsx_segment', _lib.Segment), ('vsx_half_block', _lib.HalfBlock), ('vsx_half_block_grad', _lib.HalfBlockGrad),
                       ('vsx_gemm_desc', _lib.GemmDesc), ('vsx_adamw_tensor', _lib.AdamWTensor)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), hdr, flags=re.S).group(1)
        names = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(','):
                nm = re.sub(r'\[.*?\]', '', part.strip().split()[-1]).lstrip('*')
                names.append(nm)
        assert names == [f[0] for f in cls._fields_], (cname, names, [f[0] for f in cls._fields_])


def test_random_search_candidates_fit_the_supernet():
    """Candidates drawn from the sr_tiny search space (tools/evo_eval_bench.sample_candidate: uniform choices, removed blocks propagate like
    search_utils/gen_utils.update_depth) are accepted by subnet_extents, stay inside the super-network and skip exactly the removed blocks."""
    import random
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    from vit_search_b200.evo_eval import sample_candidate
    from vit_search_b200 import supernet_config as sc
    from vit_search_b200.nets import create_model
    nd, ks = sc.network_def('sr_tiny'), sc.num_channels_to_keep('sr_tiny')
    m = create_model('flexible_vit_sr_patch14_224_patch_output', network_def=nd, num_classes=1000)
    rng = random.Random(3)
    seen_skip = False
    for _ in range(20):
        cand = sample_candidate(nd, ks, rng)
        ext = m.subnet_extents(cand)
        assert len(ext) == len(nd)
        for d, u, e in zip(cand, nd, ext):
            if d[0] == 1:
                assert e == ({'skip': True} if not d[3] else {'attn': d[1][1] * d[1][2], 'mlp': d[2][1]})
                assert d[1][1] <= u[1][1] and d[2][1] <= u[2][1]
                seen_skip |= not d[3]
            elif d[0] in (3, 4):
                assert e == {'embed': d[2] if d[0] == 3 else d[1]}
    assert seen_skip
    assert m.subnet_extents(nd)[1] == {'attn': 256, 'mlp': 768}


def test_segments_and_group_permutation_properties():
    """Host-side segment logic on random keep lists: make_segments partitions the batch into maximal runs of identical extents, and the
    group permutation is a stable permutation that makes samples of one architecture contiguous."""
    from hypothesis import given, settings, strategies as st
    from vit_search_b200 import core
    from vit_search_b200.nets.vit_sr_supernet import FlexibleDistillVisionTransformerSR as M, _runs

    keeps = st.lists(st.tuples(st.sampled_from([64, 56, 40]), st.sampled_from([64, 32, 0]), st.sampled_from([64, 0])), min_size=1, max_size=24)

    @settings(max_examples=60, deadline=None)
    @given(keeps)
    def check(rows):
        B = len(rows)
        ek, ik, ck = [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows]
        segs = core.make_segments(B, 64, ek, ik, 64, ck)
        assert segs[0].b0 == 0 and segs[-1].b1 == B and all(a.b1 == b.b0 for a, b in zip(segs, segs[1:]))
        for s in segs:
            assert all((ek[b], ik[b], ck[b]) == (s.ek, s.ik, s.ck) for b in range(s.b0, s.b1))
            assert s.active == (s.ek > 0 and s.ik > 0 and s.ck > 0)
        assert all(a.key() != b.key() for a, b in zip(segs, segs[1:]))             # runs are maximal
        runs = _runs(ek, B, 64)
        assert runs[0][0] == 0 and runs[-1][1] == B and all(k == ek[b0] for b0, _, k in runs)
        kd = [{'embed': ek}, {'attn': ik, 'layer': ck}]
        perm = M._group_permutation(kd, B)
        sig = [(ek[b], ik[b], ck[b]) for b in range(B)]
        order = list(range(B)) if perm is None else perm
        assert sorted(order) == list(range(B))
        seen, last = set(), None
        for b in order:                                                            # contiguous groups
            if sig[b] != last:
                assert sig[b] not in seen
                seen.add(sig[b])
                last = sig[b]
        for s in set(sig):                                                         # stable inside a group
            idx = [b for b in order if sig[b] == s]
            assert idx == sorted(idx)

    check()


def test_train_step_seed_sequence_resets_every_epoch():
    """engine.py:98-122: `train_iter = 0` at the top of every epoch and, for 'single' / 'hybrid' sampling, torch.manual_seed(epoch *
    10000 + train_iter) before the forward.  TrainStep must seed identically when one object is reused across epochs."""
    from vit_search_b200.engine import TrainStep

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(4))
            self.seeds = []

        def forward(self, x, patch_output_type=None):
            self.seeds.append(torch.initial_seed())
            return self.w * x.sum(), self.w * 2.0

    class Opt:
        guard = None

        def zero_grad(self):
            pass

        def step(self, **kw):
            pass

    net = Net()
    step = TrainStep(net, Opt(), criterion=lambda a, b: a.sum(), arch_sample='single')
    torch.manual_seed(99)
    before = torch.random.get_rng_state()
    for epoch, iters in ((0, 3), (1, 2), (2, 2)):
        for _ in range(iters):
            step(torch.ones(2), None, None, epoch=epoch)
    assert net.seeds == [0, 1, 2, 10000, 10001, 20000, 20001]
    assert torch.equal(torch.random.get_rng_state(), before)          # engine.py:164-165: the RNG state is restored after the draw
    step.reset_epoch(2)
    step(torch.ones(2), None, None, epoch=2)
    assert net.seeds[-1] == 20000
    # 'multi' sampling never reseeds
    net2 = Net()
    step2 = TrainStep(net2, Opt(), criterion=lambda a, b: a.sum(), arch_sample='multi')
    torch.manual_seed(5)
    step2(torch.ones(2), None, None, epoch=3)
    assert net2.seeds == [5]


def test_fused_adamw_state_dict_layout_and_round_trip():
    """main.py saves optimizer.state_dict() in every checkpoint and restores it on --resume: FusedAdamW must produce / accept the
    torch.optim.AdamW layout of timm's two parameter groups (no-decay first), so reference checkpoints interoperate."""
    from vit_search_b200.engine import FusedAdamW
    m = _small(example_per_arch=2, num_warmup_epochs=0)
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    # the same grouping built the way timm's add_weight_decay + torch.optim.AdamW does it
    decay, no_decay = [], []
    for name, p in m.named_parameters():
        (no_decay if (p.ndim <= 1 or name.endswith('.bias') or name in m.no_weight_decay()) else decay).append(p)
    ref = torch.optim.AdamW([{'params': no_decay, 'weight_decay': 0.}, {'params': decay, 'weight_decay': 0.05}], lr=1e-3)
    for p in m.parameters():
        p.grad = torch.randn_like(p) * 1e-2
    ref.step()
    ref.step()
    rsd = ref.state_dict()
    assert [g['params'] for g in opt.param_groups] == [g['params'] for g in rsd['param_groups']]
    assert [g['weight_decay'] for g in opt.param_groups] == [0.0, 0.05]
    opt.load_state_dict(rsd)                              # a reference checkpoint loads ...
    assert opt.step_count == 2
    by_index = {opt._index[n]: n for n, _, _ in opt.entries}
    params = dict(m.named_parameters())
    flat_ref = no_decay + decay
    for i, name in by_index.items():
        assert params[name] is flat_ref[i]                # ... index i means the same parameter on both sides
        assert torch.equal(opt.state[name][0], rsd['state'][i]['exp_avg']) and torch.equal(opt.state[name][1], rsd['state'][i]['exp_avg_sq'])
    out = opt.state_dict()                                # ... and what we save, torch.optim.AdamW loads
    ref2 = torch.optim.AdamW([{'params': no_decay, 'weight_decay': 0.}, {'params': decay, 'weight_decay': 0.05}], lr=1e-3)
    ref2.load_state_dict(out)
    for i in by_index:
        assert torch.equal(ref2.state_dict()['state'][i]['exp_avg'], rsd['state'][i]['exp_avg'])
        assert float(ref2.state_dict()['state'][i]['step']) == 2.0
    assert all(k in opt.param_groups[0] for k in ('lr', 'betas', 'eps', 'weight_decay'))
