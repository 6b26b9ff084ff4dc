"""Test helper: a complete synthetic model set for the detection cascade (flows + Gaussian heads).

The shipped flow pickles are absent and the shipped classifiers only make sense on the
features of those flows, so end-to-end cascade tests need a self-consistent stand-in: scenes with
face-like templates, flows fitted on windows around them (``pyfaceanalysis_b200.synthetic``), and
``mdp.nodes.GaussianClassifier``-shaped heads fitted on the flows' features with the pipeline's label
conventions (Disc: 0 = centred face ... 1 = no face; PosX / PosY in 128-px regression units; PAng in
degrees; Scale as the size ratio 0.694 .. 0.981).  Test infrastructure: uses the oracle for crops and
flow features; results are cached under build/ (which travels to the GPU box).
"""
import os
import pickle

import numpy as np

from oracle import crop as ocrop
from oracle import nodes as onodes
from pyfaceanalysis_b200 import pickles, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = (40, 20, 22.5, 0.694, 0.981, 64, 64, 128, 128)     # Pipelines/Pipeline_experimental.txt:2
HEADER_EYE = (8, 8, 0.675, 0.975, 64, 64, 64, 64)           # Pipelines/Pipeline_experimental.txt:3
NETWORK_TYPES = ["Disc1", "PosX0", "PosY0", "PAng0", "Scale0", "Disc3", "PosX1", "PosY1", "PAng1", "Scale1",
                 "Disc5", "PosX2", "PosY2", "PAng2", "Scale2", "Disc7", "Disc9"]   # stages 0..16 of the pipeline


def face_template(seed=0, size=96):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size] / float(size)
    oval = np.exp(-(((xx - 0.5) / 0.33) ** 2 + ((yy - 0.5) / 0.42) ** 2) ** 2)
    t = 120 + 90 * oval
    for (cx, cy, r, a) in ((0.34, 0.40, 0.07, -110), (0.66, 0.40, 0.07, -110), (0.5, 0.72, 0.11, -70), (0.5, 0.56, 0.04, -40)):
        t += a * np.exp(-(((xx - cx) / r) ** 2 + ((yy - cy) / (0.7 * r)) ** 2))
    t += 6 * rng.standard_normal((size, size))
    return np.clip(t, 0, 255), oval


def render_scene(H, W, faces, seed):
    """uint8 (H, W): smooth noise background with the template pasted at (cx, cy, size, angle_deg)."""
    rng = np.random.default_rng(seed)
    fy, fx = np.fft.fftfreq(H)[:, None], np.fft.fftfreq(W)[None, :]
    filt = 1.0 / (1.0 + (np.sqrt(fx * fx + fy * fy) / 0.03) ** 2)
    bg = np.fft.ifft2(np.fft.fft2(rng.standard_normal((H, W))) * filt).real
    img = 110 + 40 * bg / (bg.std() + 1e-9) + 4 * rng.standard_normal((H, W))
    tmpl, oval = face_template(0)
    ts = tmpl.shape[0]
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    for (cx, cy, size, ang) in faces:
        th = np.deg2rad(ang)
        u = ((xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)) / size + 0.5
        v = (-(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)) / size + 0.5
        inside = (u >= 0) & (u < 1) & (v >= 0) & (v < 1)
        ui = np.clip((u * ts).astype(int), 0, ts - 1)
        vi = np.clip((v * ts).astype(int), 0, ts - 1)
        alpha = np.where(inside, np.clip(oval[vi, ui] * 1.5, 0, 1), 0.0)
        img = img * (1 - alpha) + tmpl[vi, ui] * alpha
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def _window_for(face, dx, dy, dang, scale_ratio):
    """Box + angle of a window that sees `face` displaced by (dx, dy) regression pixels, rotated by dang and
    at relative size scale_ratio (0.825 = nominal)."""
    cx, cy, size, ang = face
    w = size / scale_ratio                       # the face fills `scale_ratio` of the window
    bx = cx - dx * w / 128.0
    by = cy - dy * w / 128.0
    return np.array([bx - w / 2, by - w / 2, bx + w / 2 - 1, by + w / 2 - 1]), ang + dang


def training_set(n, rng, ranges):
    """Windows around random faces with random pose errors inside `ranges` (+ pure background windows)."""
    patches, labels = [], []
    per_scene = 40
    for s in range((n + per_scene - 1) // per_scene):
        H, W = 240, 320
        face = (rng.uniform(110, 210), rng.uniform(90, 150), rng.uniform(50, 80), rng.uniform(-8, 8))
        img = render_scene(H, W, [face], int(rng.integers(1 << 30)))
        for _ in range(per_scene):
            is_face = rng.random() < 0.7
            dx = rng.uniform(-ranges["dx"], ranges["dx"])
            dy = rng.uniform(-ranges["dy"], ranges["dy"])
            da = rng.uniform(-ranges["da"], ranges["da"])
            sc = rng.uniform(0.694, 0.981)
            if is_face:
                box, ang = _window_for(face, dx, dy, da, sc)
            else:
                w = rng.uniform(40, 110)
                x0, y0 = rng.uniform(0, W - w), rng.uniform(0, H - w)
                if abs(x0 + w / 2 - face[0]) < 60 and abs(y0 + w / 2 - face[1]) < 60:
                    x0 = (x0 + 150) % (W - w)
                box, ang = np.array([x0, y0, x0 + w - 1, y0 + w - 1]), 0.0
            p = ocrop.extract_subimages(img, box[None, :], np.array([ang]))[0]
            patches.append(p.astype(np.uint8))
            labels.append((1.0 if is_face else 0.0, dx, dy, -da, sc))
    return np.asarray(patches[:n]), np.asarray(labels[:n])


def eye_training_set(n, rng):
    """Contrast-normalised 64x64 eye-box crops with known eye displacement (eye regression units: 64 = box)."""
    from oracle import controller as octl
    patches, labels = [], []
    per_scene = 30
    for s in range((n + per_scene - 1) // per_scene):
        face = (rng.uniform(110, 210), rng.uniform(90, 150), rng.uniform(55, 85), rng.uniform(-8, 8))
        img = render_scene(240, 320, [face], int(rng.integers(1 << 30)))
        w = face[2] / 0.825
        fbox = np.array([face[0] - w / 2, face[1] - w / 2, face[0] + w / 2, face[1] + w / 2])
        _, boxL, boxR = octl.eye_boxes(fbox, rot_angle=face[3])
        for _ in range(per_scene):
            eb = (boxL if rng.random() < 0.5 else boxR).copy()
            dx, dy = rng.uniform(-10, 10), rng.uniform(-10, 10)
            bw = eb[2] - eb[0]
            # displace the box so that the eye appears (dx, dy) regression pixels off-centre
            sx, sy = dx * 2.3719 * bw / 64.0 / 2.3719, dy * 2.3719 * bw / 64.0 / 2.3719
            eb[[0, 2]] += sx
            eb[[1, 3]] += sy
            p = ocrop.extract_subimages(img, eb[None, :], np.array([face[3]]))
            patches.append(ocrop.contrast_avg_std(p, 0.11, 0.15)[0])
            labels.append((dx, dy))
    return np.asarray(patches[:n]), np.asarray(labels[:n])


def fit_gaussian_classifier(feats, target, n_classes, lo, hi):
    """mdp.nodes.GaussianClassifier-shaped attribute bag: equal-width label bins, shared-ridge covariances."""
    edges = np.linspace(lo, hi, n_classes + 1)
    cls = np.clip(np.digitize(target, edges) - 1, 0, n_classes - 1)
    D = feats.shape[1]
    pooled = np.cov(feats.T) + 1e-6 * np.eye(D)
    means, inv_covs, sqrt_dets, p, labels, avg = [], [], [], [], [], []
    for c in range(n_classes):
        sel = feats[cls == c]
        if len(sel) < 3:
            continue
        cov = np.cov(sel.T) if len(sel) > D + 2 else pooled
        cov = 0.5 * cov + 0.5 * pooled
        means.append(sel.mean(axis=0))
        inv_covs.append(np.linalg.inv(cov))
        sqrt_dets.append(float(np.sqrt(np.linalg.det(cov))))
        p.append(len(sel) / float(len(feats)))
        labels.append(float(len(labels)))
        avg.append(float(target[cls == c].mean()))
    p = list(np.asarray(p) / np.sum(p))
    return pickles.new_object("mdp.nodes", "GaussianClassifier", means=means, inv_covs=inv_covs,
                              _sqrt_def_covs=sqrt_dets, p=p, labels=labels, avg_labels=np.asarray(avg),
                              _input_dim=D, _output_dim=D, input_dim=D, _dtype=np.dtype("float64"))


def build_models(seed=0, spec="S5L_64", n_train=1600):
    rng = np.random.default_rng(seed)
    wide = dict(dx=40, dy=20, da=22, )
    narrow = dict(dx=14, dy=13, da=21)
    Pw, Lw = training_set(n_train, rng, wide)
    Pn, Ln = training_set(n_train, rng, narrow)
    flows = {
        "disc_a": synthetic.make_flow(spec, seed=seed + 1, train_patches=Pn),
        "disc_b": synthetic.make_flow(spec, seed=seed + 2, train_patches=Pn),
        "pose0": synthetic.make_flow(spec, seed=seed + 3, train_patches=Pw[Lw[:, 0] > 0.5]),
        "pose1": synthetic.make_flow(spec, seed=seed + 4, train_patches=Pn[Ln[:, 0] > 0.5]),
    }

    def feats(flow, P, D):
        return onodes.flow_execute(flow, P.astype(np.float64))[:, :D]

    def disc_head(flow, P, L):
        # label 0 = centred face, 1 = far from any face (avg_labels in [0, 1] like the shipped Disc heads)
        off = np.sqrt((L[:, 1] / 40.0) ** 2 + (L[:, 2] / 20.0) ** 2) / np.sqrt(2)
        target = np.where(L[:, 0] > 0.5, 0.6 * off, 1.0)
        return fit_gaussian_classifier(feats(flow, P, 9), target, 10, 0.0, 1.0 + 1e-9)

    def pose_heads(flow, P, L, r):
        face = L[:, 0] > 0.5
        F = feats(flow, P[face], 20)
        Lf = L[face]
        return {"PosX": fit_gaussian_classifier(F[:, :10], Lf[:, 1], 25, -r["dx"], r["dx"]),
                "PosY": fit_gaussian_classifier(F[:, :10], Lf[:, 2], 25, -r["dy"], r["dy"]),
                "PAng": fit_gaussian_classifier(F, Lf[:, 3], 25, -r["da"], r["da"]),
                "Scale": fit_gaussian_classifier(F, Lf[:, 4], 25, 0.694, 0.981)}

    da, db = disc_head(flows["disc_a"], Pn, Ln), disc_head(flows["disc_b"], Pn, Ln)
    h0, h1 = pose_heads(flows["pose0"], Pw, Lw, wide), pose_heads(flows["pose1"], Pn, Ln, narrow)
    networks, classifiers = [], []
    for t in NETWORK_TYPES:
        kind, serial = t[:-1], int(t[-1])
        if kind == "Disc":
            networks.append(flows["disc_b"] if serial == 9 else flows["disc_a"])
            classifiers.append(db if serial == 9 else da)
        else:
            first = kind == "PosX"
            networks.append((flows["pose0"] if serial == 0 else flows["pose1"]) if first else None)
            classifiers.append((h0 if serial == 0 else h1)[kind])
    # eye stage: one flow, two heads (EyeLX 12 features, EyeLY 10 features like the shipped classifiers)
    Pe, Le = eye_training_set(n_train // 2, rng)
    eye_flow = synthetic.make_flow(spec, seed=seed + 5, train_patches=Pe, clip_sigmas=4.0)
    Fe = onodes.flow_execute(eye_flow, Pe)
    eye_x = fit_gaussian_classifier(Fe[:, :12], Le[:, 0], 25, -10.0, 10.0)
    eye_y = fit_gaussian_classifier(Fe[:, :10], Le[:, 1], 25, -10.0, 10.0)
    # 3 trailing placeholders (Age, Race, Gender) keep `num_networks - 5` meaningful
    types = NETWORK_TYPES + ["EyeLX", "EyeLY", "Age", "Race", "Gender"]
    return dict(header=HEADER, header_eye=HEADER_EYE, network_types=types, networks=networks + [eye_flow, eye_flow] + [None] * 3,
                classifiers=classifiers + [eye_x, eye_y] + [None] * 3, num_face_stages=len(NETWORK_TYPES))


def cached_models(seed=0, spec="S5L_64"):
    path = os.path.join(ROOT, "build", "flows", "cascade_%s_%d_v3.pckl" % (spec, seed))
    if os.path.exists(path):
        with open(path, "rb") as f:
            return pickles.loads(f.read())
    m = build_models(seed, spec)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path + ".tmp", "wb") as f:
            f.write(pickles.dumps(m))
        os.replace(path + ".tmp", path)
    except OSError:
        pass
    return m


def test_scene(seed=5, H=300, W=400, n_faces=3):
    rng = np.random.default_rng(seed)
    faces = []
    for k in range(n_faces):
        faces.append((rng.uniform(60, W - 60), rng.uniform(60, H - 60), rng.uniform(45, 85), rng.uniform(-10, 10)))
    return render_scene(H, W, faces, seed + 100), faces


test_scene.__test__ = False
