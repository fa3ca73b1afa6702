"""CPU tests of the learner-side host logic: oracle vs the reference runner goldens, RNG stream / minibatch schedule,
checkpoint format, minibatch sharding, and a world_size-2 gloo all-reduce of the flat gradient buffer."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('case', ['a', 'b', 'c'])
def test_oracle_postprocess_matches_reference_runner(case):
    from oracle.ppo_oracle import runner_postprocess, sf01, sf0
    d = np.load(os.path.join(GOLD, 'runner_%s.npz' % case))
    u, ab, g, l, rb, cb = d['params']
    rew, ret, (op, oe, ra) = runner_postprocess(int(u), int(ab), g, l, rb, cb, d['in_shaping'], d['in_main'], d['in_values'], d['in_nlp'],
                                                d['in_opp_nlp'], d['in_dones'], d['in_last_values'], d['in_last_dones'])
    assert np.array_equal(sf01(rew), d['rewards']) and np.array_equal(sf0(ra), d['ratio']) and np.array_equal(sf0(op), d['off_policy_ratio'])
    np.testing.assert_allclose(sf01(ret), d['returns'], rtol=0, atol=1.3e-7 * abs(d['returns']).max())
    assert np.array_equal(sf01(d['in_dones']), d['dones']) and np.array_equal(sf01(d['in_values']), d['values'])


def test_param_layout_and_rng_stream_match_oracle():
    from oracle import ppo_oracle as po
    from robosumo_selfplay_b200 import policies
    assert policies.param_count(121, 8) == 24529 and policies.param_shapes(121, 8) == po.param_shapes(121, 8)
    np.random.seed(42)
    a = [policies.init_params(121, 8) for _ in range(3)]
    c1 = np.random.choice(5, 1)[0]; s1 = np.arange(50); np.random.shuffle(s1)
    np.random.seed(42)
    b = [po.init_params(121, 8) for _ in range(3)]
    c2 = np.random.choice(5, 1)[0]; s2 = np.arange(50); np.random.shuffle(s2)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert c1 == c2 and np.array_equal(s1, s2)
    p = policies.unflatten_params(a[0], 121, 8)
    w = p[0].astype(np.float64)
    np.testing.assert_allclose(w.T @ w, 2.0 * np.eye(64), atol=1e-5)          # ortho_init scale sqrt(2)
    assert abs(p[8]).max() < 0.01 and (p[10] == 0).all() and (p[1] == 0).all()
    assert policies.logstd_offset(121, 8) == sum(int(np.prod(s)) for s in po.param_shapes(121, 8)[:10])


def test_reference_checkpoint_format():
    """model.ckpt of the reference (120-d obs vintage) has exactly our layout for D=120."""
    import joblib
    from robosumo_selfplay_b200 import policies
    path = '/root/reference/model.ckpt'
    if not os.path.exists(path):
        pytest.skip("reference checkout not present on this box")
    ck = joblib.load(path)
    assert [tuple(a.shape) for a in ck] == policies.param_shapes(120, 8)
    flat = policies.flatten_params(ck)
    assert flat.size == policies.param_count(120, 8) == 24401
    back = policies.unflatten_params(flat, 120, 8)
    assert all(np.array_equal(x, y) for x, y in zip(back, ck))


def test_minibatch_split_partitions_global_indices():
    from robosumo_selfplay_b200.dist import split_minibatch
    rng = np.random.RandomState(0)
    N, world = 4096, 4
    inds = rng.permutation(N)
    for s in range(0, N, 512):
        mb = inds[s:s + 512]
        parts = [split_minibatch(mb, r * N // world, (r + 1) * N // world) + r * N // world for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == sorted(mb.tolist())


def test_host_split_equals_the_per_minibatch_masks():
    """dist.host_split (the data-parallel cut made on the helper thread) against the reference-style masks of split_minibatch:
    same local indices in the same order for every minibatch, ragged last minibatch and ragged rank ranges included."""
    from robosumo_selfplay_b200.dist import host_split, split_minibatch
    rng = np.random.RandomState(3)
    N, nbt = 5000, 384                          # 13 full minibatches + one of 8
    perm = rng.permutation(N).astype(np.int32)
    bounds = [0, 1237, 2500, 2501, N]           # unequal ranges, one of a single sample
    total = np.zeros(14, np.int64)
    for r in range(4):
        lo, hi = bounds[r], bounds[r + 1]
        sp = host_split(perm, nbt, lo, hi)
        assert sp.idx.shape[0] == 14 and sp.n_total == N and sp.idx.dtype == np.int32
        for m in range(14):
            want = split_minibatch(perm[m * nbt:(m + 1) * nbt], lo, hi)
            assert sp.counts[m] == len(want) and np.array_equal(sp.idx[m, :sp.counts[m]], want)
        total += sp.counts
    assert total.tolist() == [nbt] * 13 + [8]


def _gloo_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch
    from robosumo_selfplay_b200.dist import Comm, split_minibatch
    comm = Comm(backend='gloo')
    # data-parallel minibatch: each rank sums "gradients" of its local part; the all-reduced flat buffer equals the global sum
    N = 1024
    rng = np.random.RandomState(7)
    g_all = rng.randn(N, 5)
    inds = rng.permutation(N)[:256]
    lo, hi = comm.shard(N)
    loc = split_minibatch(inds, lo, hi)
    buf = torch.tensor(np.concatenate([g_all[lo:hi][loc].sum(0), [float(len(loc))]]))
    comm.all_reduce_sum(buf)
    snap = torch.full((4,), float(rank))
    comm.broadcast(snap, 0)
    q.put((rank, buf.numpy(), g_all[inds].sum(0), snap.numpy()))


def test_gloo_world2_allreduce_and_broadcast():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, buf, want, snap in res:
        np.testing.assert_allclose(buf[:5], want, rtol=1e-12)
        assert buf[5] == 256 and (snap == 0).all()


def test_zoo_golden_matches_reference_asset():
    """tests/golden/zoo_ant_v3.npz was computed from the reference's agent-params-v3.npy: same size / checksum, and the float64
    restatement of policy_zoo/policy.py reproduces the stored outputs."""
    import hashlib
    from oracle.ppo_oracle import zoo_mlp_act
    g = np.load(os.path.join(GOLD, 'zoo_ant_v3.npz'))
    assert int(g['size']) == 24645
    path = '/root/reference/robosumo/robosumo/policy_zoo/assets/ant/mlp/agent-params-v3.npy'
    if not os.path.exists(path):
        pytest.skip("reference checkout not present on this box")
    flat = np.load(path)
    assert hashlib.sha256(flat.tobytes()).hexdigest() == str(g['sha256'])
    act, v = zoo_mlp_act(flat, g['obs'], 120, 8)
    np.testing.assert_allclose(act, g['act'], atol=1e-12); np.testing.assert_allclose(v, g['vpred'], atol=1e-9)


def test_legacy_shuffle_replay_is_bit_exact():
    """dist.legacy_shuffle == np.random.shuffle on the legacy global RandomState: same permutation, same stream afterwards
    (A10: the minibatch schedule is an integer, bit-exact target), across the MT19937 refill boundary and odd sizes."""
    from robosumo_selfplay_b200.dist import legacy_shuffle
    for seed, n in ((0, 1), (1, 2), (2, 3), (3, 1000), (4, 16384), (5, 524288), (6, 1000003)):
        np.random.seed(seed)
        np.random.normal(0, 1, (7, 5))                       # leaves a cached gaussian in the state, like ortho_init does
        a = np.arange(n)
        np.random.shuffle(a); np.random.shuffle(a)
        tail_ref = np.random.randint(0, 1 << 30, 5), np.random.normal(0, 1, 3)
        np.random.seed(seed)
        np.random.normal(0, 1, (7, 5))
        b = np.arange(n)
        legacy_shuffle(b); legacy_shuffle(b)
        tail = np.random.randint(0, 1 << 30, 5), np.random.normal(0, 1, 3)
        assert (a == b).all()
        assert (tail_ref[0] == tail[0]).all() and (tail_ref[1] == tail[1]).all()


def test_int32_shuffle_and_pipelined_cut_are_the_same_schedule():
    """EpochPermutations shuffles an int32 working array (rs_legacy_shuffle32) and, data-parallel, cuts each permutation on a second
    helper thread: the permutations must still be np.random.shuffle's, the generator must end where NumPy's would, and the cut must
    equal the masks of the un-cut permutation for every rank."""
    from robosumo_selfplay_b200.dist import EpochPermutations, split_minibatch
    n, nep, nbt, world = 30011, 3, 4096, 3
    np.random.seed(11)
    a = np.arange(n); ref = []
    for _ in range(nep):
        np.random.shuffle(a); ref.append(a.copy())
    tail_ref = np.random.randint(0, 1 << 30, 4)
    np.random.seed(11)
    got = list(EpochPermutations(n, nep, ahead=nep, dtype=np.int32))
    assert (np.random.randint(0, 1 << 30, 4) == tail_ref).all()
    for r, g in zip(ref, got):
        assert g.dtype == np.int32 and (r == g).all()
    bounds = [0, 9000, 20000, n]
    for rank in range(world):
        lo, hi = bounds[rank], bounds[rank + 1]
        np.random.seed(11)
        cuts = list(EpochPermutations(n, nep, ahead=1, dtype=np.int32, split=(nbt, lo, hi)))
        assert (np.random.randint(0, 1 << 30, 4) == tail_ref).all()
        for r, sp in zip(ref, cuts):
            for m in range((n + nbt - 1) // nbt):
                want = split_minibatch(r[m * nbt:(m + 1) * nbt], lo, hi)
                assert sp.counts[m] == len(want) and (sp.idx[m, :sp.counts[m]] == want).all()


def test_epoch_permutations_commit_only_what_is_consumed():
    """EpochPermutations hands out the same permutations as successive np.random.shuffle calls and leaves the global stream
    where the reference would be, also when the epochs stop early."""
    from robosumo_selfplay_b200.dist import EpochPermutations
    n = 50000
    np.random.seed(11)
    a = np.arange(n); ref = []
    for _ in range(3):
        np.random.shuffle(a); ref.append(a.copy())
    after3 = np.random.randint(0, 1 << 30, 4)
    np.random.seed(11)
    got = list(EpochPermutations(n, 3))
    assert len(got) == 3 and all((g == r).all() for g, r in zip(got, ref))
    assert (np.random.randint(0, 1 << 30, 4) == after3).all()
    # early stop after 2 of 6 epochs: the stream continues as if only 2 shuffles had been drawn
    np.random.seed(11)
    b = np.arange(n); np.random.shuffle(b); np.random.shuffle(b)
    after2 = np.random.randint(0, 1 << 30, 4)
    np.random.seed(11)
    it = EpochPermutations(n, 6)
    next(it); p2 = next(it); it.close()
    assert (p2 == b).all() and (np.random.randint(0, 1 << 30, 4) == after2).all()
