"""Oracle pinning, part 2: tsl::robin_map v1.0.1 iteration order (oracle/robin_table.hpp) against an independent
pure-Python emulator written from the published rules — the real tsl source is not in the reference
tree, so this is the strongest pin available ("parity unpinned" against tsl itself)."""
import numpy as np
import pytest


def _hash(k):
    x, y, z = (int(v) & 0xFFFFFFFF for v in k)
    return ((x * 73856093) & 0xFFFFFFFF ^ (y * 19349663) & 0xFFFFFFFF ^ (z * 83492791) & 0xFFFFFFFF) & ((1 << 20) - 1)


class PyRobin:
    """Straight-line emulator: power-of-two growth from 0, max load 0.5, robin-hood insert with swap-and-carry, rehash by
    re-inserting the old bucket array in order."""

    def __init__(self):
        self.b = []  # list of None | [dist, hash, value]

    def _place(self, table, h, v):
        mask = len(table) - 1
        i, d = h & mask, 0
        carry = [d, h, v]
        while True:
            if table[i] is None:
                carry[0] = d
                table[i] = carry
                return
            if d > table[i][0]:
                carry[0] = d
                table[i], carry = carry, table[i]
                d = carry[0]
            d += 1
            i = (i + 1) & mask

    def insert(self, h, v):
        n = sum(e is not None for e in self.b)
        if n >= int(len(self.b) * 0.5):
            old = self.b
            self.b = [None] * max(2, 2 * len(old))
            for e in old:
                if e is not None:
                    self._place(self.b, e[1], e[2])
        self._place(self.b, h, v)

    def order(self):
        return [e[2] for e in self.b if e is not None]


@pytest.mark.parametrize("n,spread,seed", [(1, 5, 0), (2, 5, 1), (3, 5, 2), (17, 3, 3), (100, 4, 4), (1000, 12, 5), (5000, 40, 6),
                                           (300, 2, 7), (4097, 30, 8)])
def test_iteration_order_matches_python_emulator(orc, n, spread, seed):
    rng = np.random.default_rng(seed)
    keys = np.unique(rng.integers(-spread, spread + 1, size=(4 * n, 3)), axis=0)
    rng.shuffle(keys)
    keys = keys[:n].astype(np.int32)
    order, bucket_count = orc.robin_order(keys)
    t = PyRobin()
    for i, k in enumerate(keys):
        t.insert(_hash(k), i)
    assert list(order) == t.order()
    assert bucket_count == len(t.b)
    m = len(keys)
    assert bucket_count == max(2, 1 << int(np.ceil(np.log2(max(2 * m - 1, 1)))))  # B = max(2, nextpow2(2n-1))


def test_hash_is_the_reference_20_bit_hash(orc):
    rng = np.random.default_rng(9)
    for k in rng.integers(-2 ** 20, 2 ** 20, size=(200, 3)):
        assert orc.voxel_hash(*[int(v) for v in k]) == _hash(k)
    assert orc.voxel_hash(0, 0, 0) == 0
    assert orc.voxel_hash(-1, 0, 0) == ((0xFFFFFFFF * 73856093) & 0xFFFFFFFF) & 0xFFFFF


def test_wraparound_and_collisions(orc):
    """Keys engineered to share one ideal bucket: the carried element rotates same-ideal groups (App. C)."""
    # x * 73856093 mod 2^20 is a bijection on x mod 2^20 (odd multiplier): x and x + 2^20 collide on all 20 bits
    keys = np.array([[5 + (i << 20 >> 0) % (1 << 30), 0, 0] for i in range(6)] + [[6, 0, 0], [4, 0, 0], [7, 0, 0]], dtype=np.int64)
    keys = keys.astype(np.int32)
    order, _ = orc.robin_order(keys)
    t = PyRobin()
    for i, k in enumerate(keys):
        t.insert(_hash(k), i)
    assert list(order) == t.order()


def test_voxel_downsample_is_first_point_per_voxel_in_robin_order(orc, cfg):
    """VoxelDownsample (core/Preprocessing.cpp:44-84) rebuilt from numpy + the Python emulator."""
    rng = np.random.default_rng(10)
    pts = np.c_[rng.uniform(-20, 20, (6000, 3)), rng.choice([40, 48, 50, 70, 80, 0, 10, 30, 252], 6000).astype(float)]
    for scale in (0.5, 1.5):
        out = orc.voxel_downsample(cfg, pts, scale)
        expect = []
        for g, (labels, vs) in enumerate(zip(cfg.voxel_labels, cfg.voxel_size)):
            # first-match group lookup: a label that also appears in an earlier group belongs to that one
            earlier = set(l for gl in cfg.voxel_labels[:g] for l in gl)
            sel = [i for i, p in enumerate(pts) if int(p[3]) in labels and int(p[3]) not in earlier]
            t, seen = PyRobin(), set()
            for i in sel:
                k = tuple(int(v) for v in np.trunc(pts[i, :3] / (vs * scale)))
                if k in seen:
                    continue
                seen.add(k)
                t.insert(_hash(k), i)
            expect += t.order()
        assert np.array_equal(out, pts[expect])
        assert not np.isin(out[:, 3], [30, 252]).any()  # labels in no group are dropped (:69)


class PyRobinMap(PyRobin):
    """PyRobin plus tsl's erase (backward-shift deletion) and the reference's erase-while-iterating sweep."""

    def erase_at(self, i):
        mask = len(self.b) - 1
        self.b[i] = None
        prev, cur = i, (i + 1) & mask
        while self.b[cur] is not None and self.b[cur][0] > 0:
            self.b[prev] = self.b[cur]
            self.b[prev][0] -= 1
            self.b[cur] = None
            prev, cur = cur, (cur + 1) & mask

    def sweep(self, far):
        """for (auto &[k, v] : map) if (far(v)) map.erase(k);  — the range-for's iterator is advanced AFTER the erase, from
        the erased bucket, so whatever was shifted into it is not visited (core/VoxelHashMap.cpp:176-184)."""
        i = 0
        while i < len(self.b):
            e = self.b[i]
            if e is not None and far(e[2]):
                self.erase_at(i)
            i += 1


@pytest.mark.parametrize("seed,max_distance", [(0, 12.0), (1, 6.0), (2, 25.0)])
def test_faithful_eviction_matches_python_model(orc, seed, max_distance):
    """RemovePointsFarFromLocation in the oracle's 'faithful' mode (erase while iterating a tsl::robin_map) against the
    Python table: same survivors, same iteration order afterwards; and only far voxels may survive a sweep."""
    rng = np.random.default_rng(seed)
    vs = 1.0
    pts = np.c_[rng.uniform(-30, 30, (6000, 3)) * [1, 1, 0.1], np.full(6000, 40.0)]
    m = orc.OracleMap(vs, max_distance, 20, 20, [40], evict_faithful=True)
    m.add_points(pts)
    # python model: voxels in order of first appearance, value = first point of the voxel
    t, seen = PyRobinMap(), {}
    for p in pts:
        k = tuple(int(v) for v in np.trunc(p[:3] / vs))
        if k not in seen:
            seen[k] = p[:3].copy()
            t.insert(_hash(k), k)
    keys0, _, _ = m.dump()
    assert [tuple(int(v) for v in k) for k in keys0] == t.order()  # same table before the sweep
    origin = np.array([4.0, -3.0, 0.0])
    far = lambda k: ((seen[k] - origin) ** 2).sum() > max_distance ** 2
    n_far0 = sum(far(k) for k in t.order())
    for sweep in range(2):
        m.remove_far(origin)
        t.sweep(far)
        keys, _, _ = m.dump()
        got = [tuple(int(v) for v in k) for k in keys]
        assert got == t.order(), sweep
    survivors_far = [k for k in t.order() if far(k)]
    # every erase can hide at most the one element shifted into its bucket: a sweep removes at least half of the far voxels
    assert n_far0 > 100 and len(survivors_far) <= n_far0 // 4 + 1
    clean = orc.OracleMap(vs, max_distance, 20, 20, [40], evict_faithful=False)
    clean.add_points(pts)
    clean.remove_far(origin)
    ck, _, _ = clean.dump()
    assert set(tuple(int(v) for v in k) for k in ck) == set(k for k in t.order() if not far(k)) | set()  # clean mode = exactly the near voxels
