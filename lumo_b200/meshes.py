"""Seeded synthetic stand-in meshes: the reference's scene assets (bunny, dragon,
suzanne, conference, bistro) are downloaded at run time (src/parser.rs:149-165) and are not
available offline, so the benchmark configs use deterministic procedural meshes of matching
triangle counts.  All generators return (vertices[N,3] f64, faces[F,3] i64)."""
import numpy as np


def _value_noise3(p, seed, octaves=3):
    """cheap lattice value noise on R^3, summed over octaves; deterministic in (p, seed)"""
    out = np.zeros(p.shape[0])
    amp, freq = 1.0, 1.0
    for o in range(octaves):
        q = p * freq
        i = np.floor(q).astype(np.int64); f = q - i
        f = f * f * (3 - 2 * f)

        def h(ix, iy, iz):
            n = (ix * 73856093) ^ (iy * 19349663) ^ (iz * 83492791) ^ (seed * 2654435761 + o * 97)
            n = (n ^ (n >> 13)) * 1274126177
            n = n ^ (n >> 16)
            return (n & 0xFFFFFF) / float(0xFFFFFF)
        acc = 0
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                    acc = acc + w * h(i[:, 0] + dx, i[:, 1] + dy, i[:, 2] + dz)
        out += amp * (acc - 0.5)
        amp *= 0.5; freq *= 2.0
    return out


def _grid_faces(nu, nv, wrap_u=True, wrap_v=False):
    iu = np.arange(nu if wrap_u else nu - 1); iv = np.arange(nv if wrap_v else nv - 1)
    U, V = np.meshgrid(iu, iv, indexing="ij")
    U1 = (U + 1) % nu; V1 = (V + 1) % nv
    a = U * nv + V; b = U1 * nv + V; c = U1 * nv + V1; d = U * nv + V1
    f1 = np.stack([a, b, c], -1).reshape(-1, 3); f2 = np.stack([a, c, d], -1).reshape(-1, 3)
    return np.concatenate([f1, f2]).astype(np.int64)


def displaced_sphere(n_tris, seed=7, amplitude=0.25):
    """UV sphere displaced by 3-octave value noise ("bunny" stand-in: 69 632 triangles at seed 7)."""
    nv = int(round(np.sqrt(n_tris / 4.0))) + 1
    nu = max(3, n_tris // (2 * (nv - 1)))
    u = np.arange(nu) / nu * 2 * np.pi
    v = (np.arange(nv) + 0.5) / nv * np.pi          # avoid the poles (degenerate fans)
    U, V = np.meshgrid(u, v, indexing="ij")
    d = np.stack([np.sin(V) * np.cos(U), np.cos(V), np.sin(V) * np.sin(U)], -1).reshape(-1, 3)
    r = 1.0 + amplitude * _value_noise3(d * 2.5 + 10.0, seed)
    verts = d * r[:, None]
    return verts, _grid_faces(nu, nv, True, False)


def torus_knot(n_tris, seed=11, p=2, q=3, tube=0.18, amplitude=0.04):
    """Displaced (p,q) torus-knot tube ("dragon" stand-in: 870 400 triangles at seed 11)."""
    nv = 64
    nu = max(8, n_tris // (2 * nv))
    t = np.arange(nu) / nu * 2 * np.pi

    def curve(t):
        r = 0.5 * (2 + np.cos(q * t))
        return np.stack([r * np.cos(p * t), r * np.sin(p * t), -np.sin(q * t) * 0.5], -1)
    c = curve(t); dt = 1e-4
    tan = curve(t + dt) - curve(t - dt); tan /= np.linalg.norm(tan, axis=1, keepdims=True)
    nrm = curve(t + dt) - 2 * c + curve(t - dt); nrm -= tan * np.sum(nrm * tan, axis=1, keepdims=True)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    bin_ = np.cross(tan, nrm)
    a = np.arange(nv) / nv * 2 * np.pi
    ring = np.cos(a)[None, :, None] * nrm[:, None, :] + np.sin(a)[None, :, None] * bin_[:, None, :]
    verts = (c[:, None, :] + tube * ring).reshape(-1, 3)
    verts = verts + amplitude * _value_noise3(verts * 6.0 + 5.0, seed)[:, None] * ring.reshape(-1, 3)
    return verts, _grid_faces(nu, nv, True, True)


def cube10():
    """The inline 10-triangle cube of the reference's kd-tree tests (kdtree_tests.rs:158-194) — open top."""
    v = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [1, 5, 6], [1, 6, 2], [3, 2, 6], [3, 6, 7]], dtype=np.int64)
    return v, f


def displaced_blob(nu, nv, seed, amplitude=0.2, squash=(1.0, 1.0, 1.0)):
    """Closed displaced UV sphere with exactly 2*nu*(nv-1) triangles (a chunk of the synthetic
    "conference" / "bistro" stand-ins)."""
    u = np.arange(nu) / nu * 2 * np.pi
    v = (np.arange(nv) + 0.5) / nv * np.pi
    U, V = np.meshgrid(u, v, indexing="ij")
    d = np.stack([np.sin(V) * np.cos(U), np.cos(V), np.sin(V) * np.sin(U)], -1).reshape(-1, 3)
    r = 1.0 + amplitude * _value_noise3(d * 2.0 + 3.0 * seed, seed)
    return d * r[:, None] * np.asarray(squash)[None, :], _grid_faces(nu, nv, True, False)


def height_patch(nu, nv, seed, amplitude=0.05):
    """Open height-field patch on [0,1]^2 (y up) with exactly 2*(nu-1)*(nv-1) triangles."""
    x = np.arange(nu) / (nu - 1); z = np.arange(nv) / (nv - 1)
    X, Z = np.meshgrid(x, z, indexing="ij")
    p = np.stack([X, np.zeros_like(X), Z], -1).reshape(-1, 3)
    p[:, 1] = amplitude * _value_noise3(p * 4.0 + 7.0 * seed, seed)
    return p, _grid_faces(nu, nv, False, False)
