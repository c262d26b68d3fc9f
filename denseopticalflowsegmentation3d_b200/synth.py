"""Host (numpy) twin of csrc/dofs_synth.cuh: the integer-defined synthetic traffic video of SURVEY.md
section 8d.  Bit-identical to the device generator (tests/test_synth.py checks it on the GPU); used to
feed the CPU reference the very frames the GPU path consumes, and by the CPU-only tests."""
import numpy as np

MAX_OBJECTS = 64
PERIOD = 8
_M = np.uint64(0xFFFFFFFF)


def _u32(a):
    return np.asarray(a, dtype=np.uint64) & _M


def _hash(a):
    a = _u32(a)
    a ^= a >> np.uint64(16)
    a = _u32(a * np.uint64(0x7FEB352D))
    a ^= a >> np.uint64(15)
    a = _u32(a * np.uint64(0x846CA68B))
    a ^= a >> np.uint64(16)
    return a


def _lattice(ix, iy, salt):
    return _hash(_u32(ix * np.uint64(0x9E3779B1)) ^ _hash(_u32(iy * np.uint64(0x85EBCA77)) ^ np.uint64(salt))) & np.uint64(255)


def _noise(x, y, cell, salt):
    cell = np.uint64(cell)
    ix, iy, fx, fy = x // cell, y // cell, x % cell, y % cell
    one = np.uint64(1)
    v00, v10 = _lattice(ix, iy, salt), _lattice(ix + one, iy, salt)
    v01, v11 = _lattice(ix, iy + one, salt), _lattice(ix + one, iy + one, salt)
    top = v00 * (cell - fx) + v10 * fx
    bot = v01 * (cell - fx) + v11 * fx
    return (top * (cell - fy) + bot * fy) // (cell * cell)


def _texture(x, y, s, salt, channel):
    coarse = _noise(x, y, 16 * s, salt)
    fine = _noise(x, y, 4 * s, int(_u32(salt * 3 + channel + 1)))
    return (np.uint64(2) * coarse + fine) // np.uint64(3)


def _rand(seed, i, j):
    return int(_hash(_u32(seed * 0x9E3779B1 + i * 0x85EBCA77 + j * 0xC2B2AE3D + 12345)))


def make_scene(seed, n_objects, W, H):
    """Objects of the scene: list of dict(x0, y0, w, h, dx, dy, salt) — synth_make_scene."""
    assert 0 <= n_objects <= MAX_OBJECTS
    s = max(H // 360, 1)
    m_max = 2 + 4 * s
    objs = []
    for i in range(n_objects):
        w = min(s * (40 + _rand(seed, i, 0) % 71), W // 2)
        h = min(s * (30 + _rand(seed, i, 1) % 41), H // 4)
        x_lo, x_hi = PERIOD, W - w - PERIOD
        y_lo, y_hi = H // 8 + m_max * PERIOD, H - h - m_max * PERIOD
        x0 = x_lo + _rand(seed, i, 2) % (x_hi - x_lo if x_hi > x_lo else 1)
        y0 = y_lo + _rand(seed, i, 3) % (y_hi - y_lo if y_hi > y_lo else 1)
        m = 2 + (4 * s * (y0 + h)) // H
        dy = m if (_rand(seed, i, 4) & 1) else -m
        dx = _rand(seed, i, 5) % 3 - 1
        salt = int(_hash(seed ^ _u32(0xA5A5 + i * 977)))
        objs.append(dict(x0=x0, y0=y0, w=w, h=h, dx=dx, dy=dy, salt=salt))
    return dict(seed=seed, s=s, objects=objs)


def travel(frame):
    ph = frame % (2 * PERIOD)
    return ph if ph < PERIOD else 2 * PERIOD - ph


def frame(scene, t, W, H):
    """BGR u8 frame t of the scene, [H][W][3]."""
    yy, xx = np.meshgrid(np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64), indexing="ij")
    tx, ty = xx.copy(), yy.copy()
    salt = np.full((H, W), int(_hash(scene["seed"] ^ 0xBACC0001)), dtype=np.uint64)
    done = np.zeros((H, W), bool)
    tr = travel(t)
    for o in reversed(scene["objects"]):
        ox, oy = o["x0"] + o["dx"] * tr, o["y0"] + o["dy"] * tr
        inside = (xx >= ox) & (xx < ox + o["w"]) & (yy >= oy) & (yy < oy + o["h"]) & ~done
        tx[inside] = xx[inside] - ox
        ty[inside] = yy[inside] - oy
        salt[inside] = o["salt"]
        done |= inside
    out = np.empty((H, W, 3), np.uint8)
    txu, tyu = tx.astype(np.uint64), ty.astype(np.uint64)
    for c in range(3):
        coarse = _noise(txu, tyu, 16 * scene["s"], salt)
        fine = _noise(txu, tyu, 4 * scene["s"], _u32(salt * np.uint64(3) + np.uint64(c + 1)))
        out[..., c] = ((np.uint64(2) * coarse + fine) // np.uint64(3)).astype(np.uint8)
    return out


def frames(seed, n_objects, first_frame, n_frames, W, H):
    sc = make_scene(seed, n_objects, W, H)
    return np.stack([frame(sc, first_frame + i, W, H) for i in range(n_frames)])
