"""LIMU_OPT_SPECULATE (the next scan's deskew + downsampling launch enqueued behind the current scan's registration) and the error
behaviour of limu_odom_register_*: results must not depend on hints / prefetches beyond rounding (the twist of a speculated scan is taken
by the device's log: bar 1e-9, far inside the north-star tolerance of 1e-5 m / 1e-6 rad), whatever the caller does with the slots."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="module")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def scans(pkg):
    from importlib import import_module
    synth = import_module("limu_b200.synth")
    scene = synth.Scene(seed=21)
    traj = synth.loop_trajectory(11, radius=30.0, step=0.7)
    return [synth.cast_scan(scene, traj[i], traj[i + 1], beams=16, azimuth_steps=700, seed=300 + i) for i in range(10)]


def plain_run(ctx, seq, **kw):
    k = ctx.KissICP(deskew=True, icp_max_iteration=60, speculate=False, **kw)
    out = []
    for s in seq:
        d, sr, p = k.register_frame(s)
        out.append((d, sr, p.copy(), k.stats.icp.iterations))
    dump = k.local_map().dump()
    k.close()
    return out, dump


def close_enough(got, ref):
    assert len(got) == len(ref)
    for (gd, gs, gp, gi), (rd, rs, rp, ri) in zip(got, ref):
        assert gd.shape == rd.shape and gs.shape == rs.shape and gi == ri
        np.testing.assert_allclose(gd, rd, rtol=0, atol=1e-9)
        np.testing.assert_allclose(gs, rs, rtol=0, atol=1e-9)
        np.testing.assert_allclose(gp, rp, rtol=0, atol=1e-9)


def test_hinted_device_replay_matches_plain(ctx, pkg, scans):
    import torch
    ref, ref_dump = plain_run(ctx, scans)
    staged = [torch.from_numpy(s).cuda() for s in scans]
    torch.cuda.synchronize()
    k = ctx.KissICP(deskew=True, icp_max_iteration=60, speculate=True)
    hits = 0
    for i, t in enumerate(staged):
        if i + 1 < len(staged):
            k.hint_next_dev(staged[i + 1].data_ptr(), len(scans[i + 1]))
        p = k.register_frame_dev(t.data_ptr(), len(scans[i]))
        hits += k.stats.reserved0
        np.testing.assert_allclose(p, ref[i][2], rtol=0, atol=1e-9)
        assert k.stats.icp.iterations == ref[i][3] and k.stats.n_down == len(ref[i][0]) and k.stats.n_keypoints == len(ref[i][1])
    assert hits == len(scans) - 1                       # every scan after the first was voxelized ahead of time
    dump = k.local_map().dump()
    assert np.array_equal(dump[0], ref_dump[0]) and np.array_equal(dump[1], ref_dump[1])
    np.testing.assert_allclose(dump[2], ref_dump[2], rtol=0, atol=1e-9)
    k.close()


def test_prefetched_host_replay_matches_plain(ctx, pkg, scans):
    ref, _ = plain_run(ctx, scans)
    pinned = [pkg.PinnedArray(s.shape, np.float32) for s in scans]
    for p, s in zip(pinned, scans):
        p.array[...] = s
    k = ctx.KissICP(deskew=True, icp_max_iteration=60, speculate=True)
    got, hits = [], 0
    for i in range(len(scans)):
        if i + 1 < len(scans):
            k.prefetch(pinned[i + 1].array)
        d, sr, p = k.register_frame(pinned[i].array)
        hits += k.stats.reserved0
        got.append((d, sr, p.copy(), k.stats.icp.iterations))
    assert hits == len(scans) - 1
    close_enough(got, ref)
    k.close()
    for p in pinned:
        p.free()


def test_slot_reuse_and_abandoned_prefetches(ctx, pkg, scans):
    """prefetch(B), register(A) speculates on B, prefetch(C), prefetch(D) -- D takes B's slot while the speculative launch
    may still read it, and has B's size. Registering D must give D's result, not B's."""
    eq = [s[: min(len(x) for x in scans)].copy() for s in scans]          # equal sizes: the dangerous case
    order = [0, 3, 4, 5, 2]                                               # A, D, then a few more
    ref, _ = plain_run(ctx, [eq[i] for i in order])
    pinned = [pkg.PinnedArray(s.shape, np.float32) for s in eq]
    for p, s in zip(pinned, eq):
        p.array[...] = s
    k = ctx.KissICP(deskew=True, icp_max_iteration=60, speculate=True)
    got = []
    k.prefetch(pinned[1].array)                                           # B
    d, sr, p = k.register_frame(pinned[0].array)                          # A (speculates on B)
    got.append((d, sr, p.copy(), k.stats.icp.iterations))
    k.prefetch(pinned[2].array)                                           # C
    k.prefetch(pinned[3].array)                                           # D: evicts the oldest pending upload (B)
    for j in (3, 4, 5, 2):                                                # none of them is what was speculated on last
        d, sr, p = k.register_frame(pinned[j].array)
        got.append((d, sr, p.copy(), k.stats.icp.iterations))
    close_enough(got, ref)
    k.close()
    # a hint that is not followed (device-pointer entry)
    import torch
    staged = [torch.from_numpy(s).cuda() for s in eq[:4]]
    torch.cuda.synchronize()
    ref2, _ = plain_run(ctx, [eq[0], eq[2], eq[3]])
    k = ctx.KissICP(deskew=True, icp_max_iteration=60, speculate=True)
    k.hint_next_dev(staged[1].data_ptr(), len(eq[1]))
    p0 = k.register_frame_dev(staged[0].data_ptr(), len(eq[0]))
    p2 = k.register_frame_dev(staged[2].data_ptr(), len(eq[2]))           # not the hinted scan: the speculation is discarded
    assert k.stats.reserved0 == 0
    k.set_speculate(False)
    k.hint_next_dev(staged[1].data_ptr(), len(eq[1]))                     # ignored now
    p3 = k.register_frame_dev(staged[3].data_ptr(), len(eq[3]))
    for g_, r_ in zip((p0, p2, p3), ref2):
        np.testing.assert_allclose(g_, r_[2], rtol=0, atol=1e-9)
    k.close()
    for p in pinned:
        p.free()


def test_out_of_range_point_commits_the_frame(ctx, pkg, scans):
    """a device status error (a point whose voxel index leaves the packed key range, e.g. inf) used to return before the
    pose was pushed while the map already held the scan. Now the frame is committed and the error tells the caller points were left out."""
    bad = scans[1].copy()
    bad[100, 0] = np.inf
    clean = np.delete(scans[1], 100, axis=0)
    ref, ref_dump = plain_run(ctx, [scans[0], clean, scans[2], scans[3]])
    for spec in (False, True):
        k = ctx.KissICP(deskew=True, icp_max_iteration=60, speculate=spec)
        pins = [pkg.PinnedArray(s.shape, np.float32) for s in (scans[0], bad, scans[2], scans[3])]
        for p, s in zip(pins, (scans[0], bad, scans[2], scans[3])):
            p.array[...] = s
        poses = []
        for i, p in enumerate(pins):
            if spec and i + 1 < len(pins):
                k.prefetch(pins[i + 1].array)
            if i == 1:
                with pytest.raises(pkg.LimuError) as e:
                    k.register_frame(p.array)
                assert e.value.status == -3
            else:
                poses.append(k.register_frame(p.array)[2].copy())
        allp = k.poses()
        assert len(allp) == 4                                             # the bad scan has a pose
        for i, r in enumerate(ref):
            np.testing.assert_allclose(allp[i], r[2], rtol=0, atol=1e-9)
        dump = k.local_map().dump()
        assert np.array_equal(dump[0], ref_dump[0]) and np.array_equal(dump[1], ref_dump[1])
        k.close()
        for p in pins:
            p.free()
