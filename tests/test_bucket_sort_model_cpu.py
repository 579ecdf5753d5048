"""CPU model of the rasteriser's one-pass bucket sort (csrc/raster_blend.cu, bucket_sort_tile): the exactness
argument rests on (i) the bucket map being monotone non-decreasing in the key's depth bits under float32 rounding
and (ii) the in-bucket rank by the full 64-bit key.  This restates both in numpy and checks them against np.sort
on adversarial depth distributions (narrow ranges, huge ranges, ties in depth with different ids)."""
import numpy as np
import pytest

K_BUCKETS = 1024


def _u2f_rz(d):
    """__uint2float_rz: uint32 -> float32, rounded toward zero."""
    x = d.astype(np.float64)
    f = x.astype(np.float32)
    up = f.astype(np.float64) > x
    f[up] = np.nextafter(f[up], np.float32(0.0))
    return f


def bucket_of(hi, lo, hi_max):
    scale = np.float32(K_BUCKETS) / (_u2f_rz(np.array([hi_max - lo], dtype=np.uint32))[0] + np.float32(1.0))
    b = (_u2f_rz((hi - lo).astype(np.uint32)) * scale).astype(np.float32)
    return np.minimum(K_BUCKETS - 1, b.astype(np.int64))


def bucket_sort(keys):
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    lo, mx = hi.min(), hi.max()
    b = bucket_of(hi, lo, mx)
    counts = np.bincount(b, minlength=K_BUCKETS)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    out = np.empty_like(keys)
    for i, k in enumerate(keys):                       # final place = bucket start + smaller keys of the same bucket
        same = keys[b == b[i]]
        out[start[b[i]] + int((same < k).sum())] = k
    return out, b


def _keys(depths, rng):
    ids = rng.permutation(len(depths)).astype(np.uint64)
    return (depths.astype(np.float32).view(np.uint32).astype(np.uint64) << np.uint64(32)) | ids


@pytest.mark.parametrize("case", ["scene", "narrow", "wide", "ties", "two_clusters", "single_value"])
def test_bucket_sort_model_equals_full_sort(case):
    rng = np.random.default_rng(hash(case) % 1000)
    n = 777
    depths = {"scene": rng.uniform(0.8, 1.6, n), "narrow": 1.2 + rng.uniform(0, 1e-5, n),
              "wide": np.exp(rng.uniform(np.log(0.2), np.log(1e6), n)), "ties": rng.choice([0.9, 1.0, 1.1, 1.25], n),
              "two_clusters": np.where(rng.random(n) < 0.5, 0.81, 1.59) + rng.uniform(0, 1e-3, n),
              "single_value": np.full(n, 1.2)}[case]
    keys = _keys(depths, rng)
    out, b = bucket_sort(keys)
    assert np.array_equal(out, np.sort(keys))
    # (i) monotone bucket map: sorting by key never decreases the bucket index
    order = np.argsort(keys)
    assert np.all(np.diff(b[order]) >= 0)


def test_bucket_map_monotone_on_adjacent_bit_patterns():
    """Adjacent and far-apart uint32 depth patterns around float32 rounding boundaries of (hi - lo)."""
    rng = np.random.default_rng(7)
    for _ in range(50):
        lo = np.uint32(rng.integers(0x3E000000, 0x40000000))
        span = int(rng.choice([3, 1000, 1 << 16, (1 << 24) + 5, (1 << 27) + 12345]))
        offs = np.unique(np.concatenate([rng.integers(0, span + 1, 400), np.arange(min(span + 1, 64)),
                                         np.array([span - 1, span]), (1 << 24) + np.arange(-3, 4)]))
        offs = offs[(offs >= 0) & (offs <= span)].astype(np.uint64)
        hi = (np.uint64(lo) + offs).astype(np.uint32)
        b = bucket_of(hi, lo, hi.max())
        assert np.all(np.diff(b) >= 0) and b.min() >= 0 and b.max() <= K_BUCKETS - 1
