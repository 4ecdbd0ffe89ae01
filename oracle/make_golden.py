"""ORACLE tooling: generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  ``python -m oracle.make_golden``

The reference package imports matplotlib/seaborn at import time (ysmr/__init__.py -> plot_functions.py:20-24);
they are absent here and irrelevant to the hot path, so they are replaced by MagicMock before the import
(SURVEY.md A.12).  Nothing else of the reference is touched.

Fixtures written (all small; the frames themselves are regenerated from the seeded scene, a sha256 of the bytes is
stored to detect generator drift):

  e2e_*.npz     rows of <stem>_list.csv produced by reference ``track_bacteria`` on a lossless FFV1 AVI
  link_*.npz    random detection sequences fed to the reference ``CentroidTracker`` + every row it returned
  stages_*.npz  outputs of the reference's cv2/scipy call sequence on small frames (pins cv2's CPU dispatch)
  gains.npz     GaussianSumFIR.gains for fps 30 (10/20/30 horizons)
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REFERENCE = '/root/reference'


def import_reference():
    for m in ('matplotlib', 'matplotlib.gridspec', 'matplotlib.pyplot', 'seaborn'):
        sys.modules.setdefault(m, MagicMock())
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import ysmr  # noqa: F401  (the package dir wins over /root/reference/ysmr.py)
    from ysmr import helper_file, track_eval, tracker
    return helper_file, track_eval, tracker


def reference_settings(helper_file, tmp, overrides=None):
    """Default tracking.ini written by the reference itself, then the headless/short-video overrides of SURVEY
    finding 10, parsed by the reference's own get_configs."""
    import configparser
    ini = os.path.join(tmp, 'tracking.ini')
    helper_file.create_configs(ini)
    cp = configparser.ConfigParser(allow_no_value=True)
    cp.optionxform = str
    cp.read(ini)
    base = {
        ('DISPLAY SETTINGS', 'user input'): 'False',
        ('DISPLAY SETTINGS', 'select files'): 'False',
        ('DISPLAY SETTINGS', 'display video analysis'): 'False',
        ('LOGGING SETTINGS', 'log file path'): os.path.join(tmp, 'logfile.log'),
        ('LOGGING SETTINGS', 'log to file'): 'False',
        ('BASIC TRACK DATA ANALYSIS SETTINGS', 'minimal length in seconds'): '2.0',
        ('BASIC TRACK DATA ANALYSIS SETTINGS', 'limit track length to x seconds'): '2.0',
        ('ADVANCED VIDEO SETTINGS', 'minimal frame count'): '10',
    }
    base.update(overrides or {})
    for (sec, key), val in base.items():
        cp[sec][key] = str(val)
    ini2 = os.path.join(tmp, 'tracking_mod.ini')
    with open(ini2, 'w') as fh:
        cp.write(fh)
    return helper_file.get_configs(ini2)


def write_ffv1(path, grey, fps):
    import cv2
    h, w = grey.shape[1:]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'FFV1'), fps, (w, h), True)
    assert vw.isOpened()
    for f in grey:
        vw.write(np.repeat(f[..., None], 3, axis=-1))
    vw.release()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


E2E_CASES = {
    # name: (SceneConfig kwargs, ini overrides)
    'cfg1_300': (dict(width=1228, height=922, n_frames=300, n_cells=50, seed=0), {}),
    'small_wod': (dict(width=320, height=240, n_frames=150, n_cells=14, seed=3, margin=30.0), {}),
    'small_dol': (dict(width=320, height=240, n_frames=120, n_cells=10, seed=4, margin=30.0, background=160.0,
                       intensity=80.0, noise_sigma=1.5, semi_major=2.5, semi_minor=2.5),
                  {('BASIC RECORDING SETTINGS', 'white bacteria on dark background'): 'False',
                   ('BASIC RECORDING SETTINGS', 'rod shaped bacteria'): 'False'}),
    'small_single': (dict(width=200, height=160, n_frames=90, n_cells=8, seed=5, margin=30.0),
                     {('ADVANCED VIDEO SETTINGS', 'adaptive double threshold'): '0.0'}),
    'small_meanstd': (dict(width=200, height=160, n_frames=200, n_cells=8, seed=6, margin=30.0),
                      {('ADVANCED VIDEO SETTINGS', 'adaptive double threshold'): '-1.0',
                       ('BASIC RECORDING SETTINGS', 'threshold offset for detection'): '25'}),
}


def make_e2e(helper_file, track_eval):
    from ysmr_b200.synth import SceneConfig, make_scene, render_frames
    for name, (skw, ini) in E2E_CASES.items():
        cfg = SceneConfig(**skw)
        scene = make_scene(cfg)
        grey = render_frames(scene)
        with tempfile.TemporaryDirectory() as tmp:
            settings = reference_settings(helper_file, tmp, ini)
            assert settings is not None
            video = os.path.join(tmp, f'{name}.avi')
            write_ffv1(video, grey, cfg.fps)
            out = os.path.join(tmp, 'res'); os.makedirs(out)
            res = track_eval.track_bacteria(video, settings, out)
            assert res is not None, name
            df, fps, fh, fw, csv = res
            with open(csv) as fhh:
                csv_head = ''.join(fhh.readlines()[:6])
        rows = df[['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']] \
            .to_numpy(np.float64)
        keys = {f'scene_{k}': v for k, v in skw.items()}
        np.savez_compressed(
            os.path.join(GOLDEN, f'e2e_{name}.npz'), rows=rows, fps=fps, frame_height=fh, frame_width=fw,
            frames_sha256=sha(grey), csv_head=csv_head,
            white_on_dark=settings['white bacteria on dark background'],
            # NB the reference flipped the sign in place (track_eval.py:132); store the ini value
            offset=abs(settings['threshold offset for detection']),
            adt=settings['adaptive double threshold'], **keys)
        print('e2e', name, rows.shape, 'ids', int(rows[:, 0].max()) + 1)


def random_detection_sequence(rng, n_frames, n_cells, width=400.0, height=300.0, p_miss=0.08, p_birth=0.04,
                              p_death=0.01, p_empty=0.02, burst_every=37):
    """Moving points with dropouts, births (several per frame now and then), deaths and empty frames."""
    pts = rng.uniform(20, [width - 20, height - 20], (n_cells, 2))
    vel = rng.normal(0, 1.0, (n_cells, 2))
    seq = []
    for t in range(n_frames):
        vel = 0.9 * vel + rng.normal(0, 0.4, vel.shape)
        pts = pts + vel
        alive = rng.random(len(pts)) > p_death
        pts, vel = pts[alive], vel[alive]
        births = rng.poisson(p_birth * 5)
        if burst_every and t % burst_every == burst_every - 1:
            births += rng.integers(2, 9)
        if births:
            pts = np.vstack([pts, rng.uniform(20, [width - 20, height - 20], (births, 2))])
            vel = np.vstack([vel, rng.normal(0, 1.0, (births, 2))])
        if rng.random() < p_empty:
            seq.append(np.zeros((0, 5), np.float32)); continue
        seen = rng.random(len(pts)) > p_miss
        det = pts[seen] + rng.normal(0, 0.15, (int(seen.sum()), 2))
        order = rng.permutation(len(det))
        rec = np.zeros((len(det), 5), np.float32)
        rec[:, :2] = det[order]
        rec[:, 2] = rng.uniform(2, 9, len(det)); rec[:, 3] = rng.uniform(1, 4, len(det))
        rec[:, 4] = rng.uniform(-90, 0, len(det))
        seq.append(rec)
    return seq


LINK_CASES = {
    'a': dict(seed=1, n_frames=400, n_cells=12),
    'b': dict(seed=2, n_frames=300, n_cells=40, p_miss=0.15, p_birth=0.1),
    'c_dense': dict(seed=3, n_frames=60, n_cells=300, width=1228.0, height=922.0, p_miss=0.05),
    'd_sparse': dict(seed=4, n_frames=250, n_cells=2, p_miss=0.3, p_empty=0.2, p_birth=0.02, burst_every=0),
}


def run_reference_tracker(tracker, seq, fps, use_gsff=True):
    ct = tracker.CentroidTracker(max_disappeared=fps, use_gsff=use_gsff, fps=fps, n_min=0, n_max=30, n_f=3)
    rows = []
    for t, rec in enumerate(seq):
        rects = [((float(r[0]), float(r[1])), (float(r[2]), float(r[3]), float(r[4]))) for r in rec]
        objects, info = ct.update(rects)
        for oid, xy in objects.items():
            w, h, d = info[oid]
            rows.append((t, oid, xy[0], xy[1], w, h, d))
    return np.array(rows, np.float64).reshape(-1, 7)


def make_link(tracker):
    for name, kw in LINK_CASES.items():
        kw = dict(kw)
        rng = np.random.default_rng(kw.pop('seed'))
        seq = random_detection_sequence(rng, **kw)
        counts = np.array([len(s) for s in seq], np.int32)
        flat = np.concatenate(seq, axis=0) if len(seq) else np.zeros((0, 5), np.float32)
        for gs in (True, False):
            rows = run_reference_tracker(tracker, seq, 30.0, use_gsff=gs)
            np.savez_compressed(os.path.join(GOLDEN, f'link_{name}_{"gsff" if gs else "raw"}.npz'),
                                counts=counts, dets=flat, rows=rows, fps=30.0, use_gsff=gs)
            print('link', name, gs, rows.shape, 'ids', int(rows[:, 1].max()) + 1 if len(rows) else 0)


def make_stages():
    """cv2/scipy outputs on small frames in THIS container (AVX2 dispatch on an AVX-512 host)."""
    from oracle import ref_stages
    from ysmr_b200.synth import SceneConfig, make_scene, render_frames
    cases = {
        'wod': (SceneConfig(width=164, height=120, n_frames=3, n_cells=10, seed=11, margin=20.0),
                ref_stages.DetectSettings(True, 5, 2.0)),
        'dol': (SceneConfig(width=168, height=120, n_frames=3, n_cells=8, seed=12, margin=20.0, background=160.0,
                            intensity=80.0, noise_sigma=1.5, semi_major=2.5, semi_minor=2.5),
                ref_stages.DetectSettings(False, 5, 2.0)),
        'odd': (SceneConfig(width=157, height=99, n_frames=2, n_cells=8, seed=13, margin=20.0),
                ref_stages.DetectSettings(True, 4, 1.5)),
    }
    for name, (cfg, st) in cases.items():
        grey = render_frames(make_scene(cfg))
        out = {'grey': grey, 'white_on_dark': st.white_on_dark, 'offset': st.offset, 'adt': st.adt}
        rng = np.random.default_rng(99)
        bgr = rng.integers(0, 256, grey.shape + (3,), dtype=np.uint8)       # true colour input for cvtColor
        import cv2
        out['bgr'] = bgr
        out['bgr_gray'] = np.stack([cv2.cvtColor(b, cv2.COLOR_BGR2GRAY) for b in bgr])
        for key in ('blurred', 'mask', 'markers', 'out'):
            out[key] = []
        rects, counts = [], []
        for f in grey:
            r = ref_stages.detect_frame(f, st)
            for key in ('blurred', 'mask', 'markers', 'out'):
                out[key].append(r[key])
            a = ref_stages.rects_to_array(r['rects'])
            rects.append(a); counts.append(len(a))
        for key in ('blurred', 'mask', 'markers', 'out'):
            out[key] = np.stack(out[key])
        out['rects'] = np.concatenate(rects); out['counts'] = np.array(counts, np.int32)
        np.savez_compressed(os.path.join(GOLDEN, f'stages_{name}.npz'), **out)
        print('stages', name, counts)


def make_gains(tracker):
    ct = tracker.CentroidTracker(max_disappeared=30.0, fps=30.0, n_min=0, n_max=30, n_f=3)
    np.savez_compressed(os.path.join(GOLDEN, 'gains.npz'), g10=ct.gsff.gains[0], g20=ct.gsff.gains[1],
                        g30=ct.gsff.gains[2], n_i=np.array(ct.gsff.n_i))
    print('gains', [g.shape for g in ct.gsff.gains])


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    sys.path.insert(0, ROOT)
    helper_file, track_eval, tracker = import_reference()
    which = sys.argv[1:] or ['gains', 'link', 'stages', 'e2e']
    if 'gains' in which:
        make_gains(tracker)
    if 'link' in which:
        make_link(tracker)
    if 'stages' in which:
        make_stages()
    if 'e2e' in which:
        make_e2e(helper_file, track_eval)


if __name__ == '__main__':
    main()
