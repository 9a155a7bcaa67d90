"""CPU emulation of the empty-space trimming (csrc/capi.cu: cell flags -> brick occupancy -> Chebyshev distance field;
csrc/common.cuh: first_occupied_brick / occupied_alpha_range; csrc/trilinear.cu: trim_sample_range), vectorised over
rays in fp32, against brute force: on random sparse volumes with odd shapes and random rays (sources far away, close by
and INSIDE the volume; axis-parallel and grazing directions), NO sample that has a non-zero voxel among its 8 corners may
fall outside the trimmed sample range [kb, ke) -- the property the bit-identity of images and Jacobians rests on.
Also prints how tight the trimming is (samples marched / samples that touch a non-zero voxel).
    PYTHONPATH=. python scripts/emulate_trim_walk.py
"""
import numpy as np

f32 = np.float32
BK, CAP = 8, 24


def distance_field(vol):
    D = np.array(vol.shape)
    m = (D + 1) // 2
    pad = np.zeros(m * 2, bool)
    pad[:D[0], :D[1], :D[2]] = vol != 0
    cells = pad.reshape(m[0], 2, m[1], 2, m[2], 2).any(axis=(1, 3, 5))
    nb = (D + BK - 1) // BK
    H = BK // 2
    occ = np.zeros(nb, bool)
    for bx in range(nb[0]):
        for by in range(nb[1]):
            for bz in range(nb[2]):
                occ[bx, by, bz] = cells[max(bx * H - 1, 0):min((bx + 1) * H + 1, m[0]),
                                        max(by * H - 1, 0):min((by + 1) * H + 1, m[1]),
                                        max(bz * H - 1, 0):min((bz + 1) * H + 1, m[2])].any()
    dist = np.where(occ, 0, CAP).astype(np.int32)
    for axis in (2, 1, 0):  # out(c) = min_o max(|o|, in(c + o e_axis))
        out = dist.copy()
        n = dist.shape[axis]
        for o in range(1, CAP):
            for sgn in (-1, 1):
                sh = np.full_like(dist, CAP)
                src = [slice(None)] * 3
                dst = [slice(None)] * 3
                if sgn < 0:
                    src[axis], dst[axis] = slice(0, n - o), slice(o, n)
                else:
                    src[axis], dst[axis] = slice(o, n), slice(0, n - o)
                if n - o > 0:
                    sh[tuple(dst)] = dist[tuple(src)]
                out = np.minimum(out, np.maximum(o, sh))
        dist = out
    # cross-check the separable transform against the definition on a few bricks
    oc = np.argwhere(occ)
    rng = np.random.default_rng(0)
    for c in rng.integers(0, nb, size=(20, 3)):
        want = min(CAP, int(np.abs(oc - c).max(axis=1).min())) if len(oc) else CAP
        assert dist[tuple(c)] == want, (c, dist[tuple(c)], want)
    return dist.astype(np.uint8), nb


def first_occupied(dist, nb, s, d, t0, t1):
    """first_occupied_brick, all rays at once; returns entry t (inf: none)."""
    n = len(s)
    moving = d != 0
    with np.errstate(divide="ignore"):
        invd = np.where(moving, f32(1) / np.where(moving, d, f32(1)), f32(0)).astype(f32)
    off = np.where(moving, f32(0), f32(np.inf)).astype(f32)
    up = (d > 0).astype(f32)
    dmax = np.abs(d).max(axis=1)
    res = np.full(n, np.inf, f32)
    done = ~(dmax > 0)
    res[done] = t0[done]
    with np.errstate(divide="ignore"):
        jump = (f32(BK) / dmax).astype(f32)
    probe = (f32(0.01) * jump).astype(f32)
    t = t0.copy()
    entered = t0.copy()
    for _ in range(4 * CAP + 64):
        if done.all():
            break
        cf = np.floor((t[:, None] * d + s).astype(f32) * f32(1.0 / BK)).astype(f32)
        c = np.clip(cf.astype(np.int64), 0, nb - 1)
        Dd = dist[c[:, 0], c[:, 1], c[:, 2]].astype(np.int32)
        hit = ~done & (Dd == 0)
        res[hit] = entered[hit]
        done |= hit
        with np.errstate(invalid="ignore"):
            t_exit = ((((cf + up) * f32(BK)).astype(f32) - s) * invd + off).astype(f32).min(axis=1)
        entered_new = np.maximum(t_exit, t)
        t_new = np.maximum(entered_new + probe, ((Dd - 1).astype(f32) * jump + t).astype(f32))
        finished = ~done & ~(entered_new < t1)
        done |= finished  # res stays inf
        t_new = np.minimum(t_new, t1)
        upd = ~done
        entered[upd] = entered_new[upd]
        t[upd] = t_new[upd]
    res[~done] = entered[~done]  # step budget exhausted: conservative
    return res


def trimmed_range(vol, dist, nb, s, d, npts):
    """alpha_range + trim_sample_range for every ray: (amin, span, kb, ke)."""
    D = np.array(vol.shape, f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        a0 = ((f32(0) - s) / d).astype(f32)
        a1 = (((D - 1) - s) / d).astype(f32)
    amin = np.clip(np.nanmax(np.minimum(a0, a1), axis=1), 0, None).astype(f32)
    amax = np.clip(np.nanmin(np.maximum(a0, a1), axis=1), None, 1).astype(f32)
    span = (amax - amin).astype(f32)
    kb = np.zeros(len(s), np.int64)
    ke = np.full(len(s), npts, np.int64)
    nz = np.argwhere(vol != 0)
    lo, hi = (nz.min(0) - 2).astype(f32), (nz.max(0) + 2).astype(f32)
    ok = span > 0
    lstep = f32(1) / f32(npts - 1)
    first = amin.copy()
    last = (lstep * f32(npts - 1) * span + amin).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        b0, b1 = ((lo - s) / d).astype(f32), ((hi - s) / d).astype(f32)
    still = (d == 0) & ~((s > lo) & (s < hi))
    tin = np.maximum(first, np.nanmax(np.where(d != 0, np.minimum(b0, b1), -np.inf), axis=1)).astype(f32)
    tout = np.minimum(last, np.nanmin(np.where(d != 0, np.maximum(b0, b1), np.inf), axis=1)).astype(f32)
    none = still.any(axis=1) | ~(tin <= tout)
    f = first_occupied(dist, nb, s, d, tin, tout)
    none |= np.isinf(f)
    lb = first_occupied(dist, nb, s, -d, -tout, -tin)
    l = np.where(np.isinf(lb), tout, -lb).astype(f32)
    f, l = np.minimum(f, l), np.maximum(f, l)
    with np.errstate(divide="ignore", invalid="ignore"):
        sc = (f32(npts - 1) / span).astype(f32)
        kin = np.clip((f - amin) * sc, -4, npts + 4)
        kout = np.clip((l - amin) * sc, -4, npts + 4)
    kb2 = np.maximum(kb, np.floor(kin) - 2)
    ke2 = np.minimum(ke, np.ceil(kout) + 3)
    ke2 = np.maximum(ke2, kb2)
    kb = np.where(ok & ~none, kb2, kb).astype(np.int64)
    ke = np.where(ok & ~none, ke2, np.where(ok & none, kb, ke)).astype(np.int64)
    return amin, span, kb, ke


def main():
    rng = np.random.default_rng(7)
    npts = 160
    shapes = [(50, 45, 61), (33, 64, 47), (71, 39, 58), (17, 90, 23), (64, 64, 64), (24, 24, 25)]
    tot_bad = tot_rays = 0
    marched = touching = 0
    for shape in shapes:
        vol = np.zeros(shape, f32)
        for _ in range(rng.integers(1, 6)):
            lo = [rng.integers(0, n) for n in shape]
            ext = [rng.integers(1, max(2, n // 3)) for n in shape]
            vol[tuple(slice(a, min(a + e, n)) for a, e, n in zip(lo, ext, shape))] = 1.0
        dist, nb = distance_field(vol)
        D = np.array(shape, f32)
        n = 3000
        centre = (D - 1) / 2
        # sources: far (C-arm like), near, inside the volume; targets on the far side; some axis-parallel / grazing
        kind = rng.integers(0, 4, n)
        dirs = rng.normal(size=(n, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        radius = np.where(kind == 0, 6.0, np.where(kind == 1, 1.2, 0.3))[:, None] * D.max()
        s = centre + dirs * radius + rng.normal(size=(n, 3)) * 2
        s[kind == 3] = rng.uniform(0, 1, size=((kind == 3).sum(), 3)) * (D - 1)  # inside
        t = centre - dirs * D.max() * 2 + rng.normal(size=(n, 3)) * D.max() * 0.6
        axis_par = rng.random(n) < 0.08
        t[axis_par, 1:] = s[axis_par, 1:]  # parallel to axis 0 (d = 0 on two axes before eps)
        graze = rng.random(n) < 0.08
        s[graze, 1] = np.round(s[graze, 1])  # sources on integer planes
        s, t = s.astype(f32), t.astype(f32)
        d = ((t - s).astype(f32) + f32(1e-8)).astype(f32)
        amin, span, kb, ke = trimmed_range(vol, dist, nb, s, d, npts)
        # brute force: which samples touch a non-zero voxel?
        u = np.linspace(0, 1, npts, dtype=f32)
        alpha = (u[None, :] * span[:, None] + amin[:, None]).astype(f32)
        pos = (alpha[..., None] * d[:, None, :] + s[:, None, :]).astype(f32)
        base = np.floor(pos).astype(np.int64)
        touch = np.zeros(pos.shape[:2], bool)
        for ox in (0, 1):
            for oy in (0, 1):
                for oz in (0, 1):
                    c = base + np.array([ox, oy, oz])
                    inside = ((c >= 0) & (c < np.array(shape))).all(-1)
                    cc = np.clip(c, 0, np.array(shape) - 1)
                    touch |= inside & (vol[cc[..., 0], cc[..., 1], cc[..., 2]] != 0)
        k = np.arange(npts)[None, :]
        kept = (k >= kb[:, None]) & (k < ke[:, None])
        valid = (span > 0)[:, None]  # the kernel leaves backward / empty ranges alone (marches them in full)
        bad = (touch & ~kept & valid).sum()
        tot_bad += int(bad)
        tot_rays += n
        marched += int((kept & valid).sum())
        touching += int((touch & valid).sum())
        print(f"shape {shape}: {n} rays, samples touching density {int((touch & valid).sum())}, marched "
              f"{int((kept & valid).sum())}, missed {int(bad)}")
    print(f"rays {tot_rays} missed samples {tot_bad} marched/touching {marched / max(1, touching):.3f}")
    assert tot_bad == 0


if __name__ == "__main__":
    main()
