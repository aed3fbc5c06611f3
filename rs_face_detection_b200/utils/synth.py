"""Synthetic workloads of the exact RetinaFace shapes (no network, no model server offline): SURVEY.md §8(d).

Plain numpy on the host; bench.py moves the arrays to the device.  Used by tests/ and bench.py only.
"""
import numpy as np

ARCFACE_TEMPLATE = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                             [41.5493, 92.3655], [70.7299, 92.2041]], np.float32)  # config.rs:46-52

# base anchors of the RetinaFace config (face_detection.rs:55-98), [stride][a][4]
BASE_ANCHORS = np.array([[[-248, -248, 263, 263], [-120, -120, 135, 135]],
                         [[-56, -56, 71, 71], [-24, -24, 39, 39]],
                         [[-8, -8, 23, 23], [0, 0, 15, 15]]], np.float32)
STRIDES = (32, 16, 8)


def _anchor_planes(image_size=640):
    planes = []
    for s, stride in enumerate(STRIDES):
        n = (image_size + stride - 1) // stride
        ys, xs = np.meshgrid(np.arange(n, dtype=np.float32) * stride, np.arange(n, dtype=np.float32) * stride, indexing="ij")
        shift = np.stack([xs, ys, xs, ys], -1)[:, :, None, :]           # (H,W,1,4)
        planes.append(BASE_ANCHORS[s][None, None] + shift)              # (H,W,A,4)
    return planes


def make_heads(B, seed=1234, n_faces=20, image_size=640, content_hw=(640, 640), bg=True):
    """9 head tensors (B,C,H,W) f32 in net_out order [score,bbox,lmk] x strides 32,16,8 plus the planted boxes.

    Scores are a softmax pair (bg+fg=1); fg ~ Beta(0.2,4) background plus `n_faces` planted faces per image, each
    lighting the anchors whose IoU with it exceeds 0.35 with fg ~ U(0.75,0.999), regression deltas = true target +
    N(0,0.05) and landmark deltas pointing at a template-shaped 5-point set inside the face.
    """
    rng = np.random.default_rng(seed)
    planes = _anchor_planes(image_size)
    A = 2
    heads = []
    for s in range(3):
        n = planes[s].shape[0]
        fg = rng.beta(0.2, 4.0, (B, A, n, n)).astype(np.float32) if bg else np.zeros((B, A, n, n), np.float32)
        fg = np.minimum(fg, np.float32(0.65))
        bbox = np.empty((B, 4 * A, n, n), np.float32)
        bbox[:, 0::4] = rng.normal(0, 0.3, (B, A, n, n))
        bbox[:, 1::4] = rng.normal(0, 0.3, (B, A, n, n))
        bbox[:, 2::4] = rng.normal(0, 0.2, (B, A, n, n))
        bbox[:, 3::4] = rng.normal(0, 0.2, (B, A, n, n))
        lmk = rng.normal(0, 0.3, (B, 10 * A, n, n)).astype(np.float32)
        heads.append([fg, bbox, lmk])
    ch, cw = content_hw
    faces = np.empty((B, n_faces, 4), np.float32)
    tn = (ARCFACE_TEMPLATE - 56.0) / 112.0
    for b in range(B):
        side = rng.uniform(24, min(300, 0.8 * min(ch, cw)), n_faces)
        cx = rng.uniform(side / 2, cw - side / 2)
        cy = rng.uniform(side / 2, ch - side / 2)
        gt = np.stack([cx - side / 2, cy - side / 2, cx + side / 2, cy + side / 2], 1).astype(np.float32)
        faces[b] = gt
        for s in range(3):
            an = planes[s].reshape(-1, 4)                               # (H*W*A,4) order (h,w,a)
            n = planes[s].shape[0]
            fg, bbox, lmk = heads[s]
            aw = an[:, 2] - an[:, 0] + 1
            ah = an[:, 3] - an[:, 1] + 1
            acx = an[:, 0] + 0.5 * (aw - 1)
            acy = an[:, 1] + 0.5 * (ah - 1)
            for f in range(n_faces):
                g = gt[f]
                iw = np.minimum(an[:, 2], g[2]) - np.maximum(an[:, 0], g[0]) + 1
                ih = np.minimum(an[:, 3], g[3]) - np.maximum(an[:, 1], g[1]) + 1
                inter = np.clip(iw, 0, None) * np.clip(ih, 0, None)
                gw, gh = g[2] - g[0] + 1, g[3] - g[1] + 1
                iou = inter / (aw * ah + gw * gh - inter)
                hit = np.nonzero(iou > 0.35)[0]
                if len(hit) == 0:
                    continue
                hw_idx, a_idx = hit // A, hit % A
                hh, ww = hw_idx // n, hw_idx % n
                gcx, gcy = g[0] + 0.5 * (gw - 1), g[1] + 0.5 * (gh - 1)
                fg[b, a_idx, hh, ww] = rng.uniform(0.75, 0.999, len(hit))
                bbox[b, 4 * a_idx + 0, hh, ww] = (gcx - acx[hit]) / aw[hit] + rng.normal(0, 0.05, len(hit))
                bbox[b, 4 * a_idx + 1, hh, ww] = (gcy - acy[hit]) / ah[hit] + rng.normal(0, 0.05, len(hit))
                bbox[b, 4 * a_idx + 2, hh, ww] = np.log(gw / aw[hit]) + rng.normal(0, 0.05, len(hit))
                bbox[b, 4 * a_idx + 3, hh, ww] = np.log(gh / ah[hit]) + rng.normal(0, 0.05, len(hit))
                for p in range(5):
                    px = gcx + tn[p, 0] * gw + rng.normal(0, 0.01 * gw, len(hit))
                    py = gcy + tn[p, 1] * gh + rng.normal(0, 0.01 * gh, len(hit))
                    lmk[b, 10 * a_idx + 2 * p, hh, ww] = (px - acx[hit]) / aw[hit]
                    lmk[b, 10 * a_idx + 2 * p + 1, hh, ww] = (py - acy[hit]) / ah[hit]
    out = []
    for s in range(3):
        fg, bbox, lmk = heads[s]
        scores = np.concatenate([1.0 - fg, fg], 1).astype(np.float32)  # (B,2A,H,W): bg channels first (face_detection.rs:322)
        out += [np.ascontiguousarray(scores), np.ascontiguousarray(bbox.astype(np.float32)), np.ascontiguousarray(lmk)]
    return out, faces


def make_dense_heads(B, seed=7, image_size=640):
    """Every anchor scores above 0.02 (the stress threshold): K = all anchors per image."""
    rng = np.random.default_rng(seed)
    out = []
    for stride in STRIDES:
        n = (image_size + stride - 1) // stride
        fg = rng.uniform(0.02, 1.0, (B, 2, n, n)).astype(np.float32)
        fg = np.maximum(fg, np.float32(0.02))
        scores = np.concatenate([1 - fg, fg], 1).astype(np.float32)
        bbox = rng.normal(0, 0.3, (B, 8, n, n)).astype(np.float32)
        lmk = rng.normal(0, 0.3, (B, 20, n, n)).astype(np.float32)
        out += [scores, bbox, lmk]
    return out


def make_frame(h, w, seed):
    """BGR u8 frame: low-frequency gradient + U{0..255} noise (noise makes 1-LSB errors visible)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    base = np.stack([127 + 100 * np.sin(xx / (0.13 * w) + c) * np.cos(yy / (0.21 * h) - c) for c in range(3)], -1)
    noise = rng.integers(0, 256, (h, w, 3)).astype(np.float32)
    return np.clip(0.6 * base + 0.4 * noise, 0, 255).astype(np.uint8)


def make_crowd_boxes(N=100000, seed=42, canvas=(3840, 2160), n_faces=5000, dup_frac=0.01):
    """C3 dense-crowd NMS stress: n_faces ground-truth faces x N/n_faces jittered candidates, scores U(0.02,1),
    ~1% exact score duplicates to exercise the stable ordering.  Returns dets (N,5) f32."""
    rng = np.random.default_rng(seed)
    per = max(1, N // n_faces)
    side = rng.uniform(12, 64, n_faces)
    cx = rng.uniform(0, canvas[0], n_faces)
    cy = rng.uniform(0, canvas[1], n_faces)
    idx = np.arange(N) % n_faces if per * n_faces != N else np.repeat(np.arange(n_faces), per)
    s = side[idx] * (1 + rng.normal(0, 0.1, N))
    s = np.clip(s, 4, None)
    x = cx[idx] + rng.normal(0, 0.1, N) * side[idx]
    y = cy[idx] + rng.normal(0, 0.1, N) * side[idx]
    sc = rng.uniform(0.02, 1.0, N).astype(np.float32)
    ndup = int(N * dup_frac)
    if ndup > 0:
        src = rng.integers(0, N, ndup)
        dst = rng.integers(0, N, ndup)
        sc[dst] = sc[src]
    dets = np.stack([x - s / 2, y - s / 2, x + s / 2, y + s / 2, sc], 1).astype(np.float32)
    perm = rng.permutation(N)
    return np.ascontiguousarray(dets[perm])


def make_landmarks(F, seed, frame_hw=(1080, 1920), outside_frac=0.05):
    """C4: the ArcFace template under a random similarity (scale 0.4-3.0, rotation +-35 deg) + N(0, 0.01*side) noise."""
    rng = np.random.default_rng(seed)
    h, w = frame_hw
    s = rng.uniform(0.4, 3.0, F)
    th = np.deg2rad(rng.uniform(-35, 35, F))
    c, sn = np.cos(th) * s, np.sin(th) * s
    t = (ARCFACE_TEMPLATE - 56.0)[None]                                  # (1,5,2)
    x = c[:, None] * t[..., 0] - sn[:, None] * t[..., 1]
    y = sn[:, None] * t[..., 0] + c[:, None] * t[..., 1]
    cx = rng.uniform(60, w - 60, F)
    cy = rng.uniform(60, h - 60, F)
    out_mask = rng.uniform(0, 1, F) < outside_frac
    cx[out_mask] = rng.choice([-20.0, w + 20.0], out_mask.sum())
    side = 112 * s
    pts = np.stack([x + cx[:, None], y + cy[:, None]], -1) + rng.normal(0, 1, (F, 5, 2)) * (0.01 * side)[:, None, None]
    return pts.astype(np.float32)
