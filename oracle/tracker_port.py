"""ORACLE (test infrastructure only -- never imported by the product path ``ysmr_b200/``).

CPU restatement of YSMR's linker: ``CentroidTracker`` (/root/reference/ysmr/tracker.py:27-230) and the in-loop
Gaussian-sum FIR filter ``GaussianSumFIR`` (/root/reference/ysmr/gsff.py:28-347).  Written from the algorithm
(SURVEY.md A.9-A.11), array based, and pinned against the *imported reference itself* by
``oracle/make_golden.py`` (fixtures in ``tests/golden/link_*.npz``; checked by tests/test_oracle_golden.py).

It deliberately keeps the reference's numerical library calls (``scipy.spatial.distance.cdist``,
``numpy.dot`` on the gain matrices, numpy's default argsort) so that it is also a fair stand-in for the
reference's CPU cost when ``bench.py`` times the CPU baseline on a box where /root/reference is absent.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.distance import cdist

from .setorder import unused_cols_order

LIKELIHOOD_MIN = 10 ** -20          # tracker.py:67


def horizon_sizes(n_min, n_max, n_f):
    """gsff.py:103-109 (equation 17): n_i = int(n_min + (n_max - n_min) / n_f * i), i = 1..n_f."""
    step = (n_max - n_min) / n_f
    return [int(n_min + step * i) for i in range(1, n_f + 1)]


def lsf_gain(n, dt):
    """gsff.py:112-153 (equations 13, 14): K = (L^T L)^-1 L^T with L = H_bar A^-n, same numpy calls and
    operation order as the reference so the float64 gains are bit-identical."""
    a = np.array([[1, 0, dt, 0], [0, 1, 0, dt], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    c = np.array([[1, 0, 0, 0], [0, 1, 0, 0]])
    h_bar = c
    a_n = a
    for _ in range(n - 1):
        h_bar = np.concatenate((h_bar, np.dot(c, a_n)), axis=0)
        a_n = np.dot(a_n, a)
    l_bar = np.dot(h_bar, np.linalg.matrix_power(np.linalg.inv(a), n))
    return np.dot(np.linalg.inv(np.dot(l_bar.T, l_bar)), l_bar.T)


class GsffBank:
    """Stateless filter bank; per-track state lives in a small dict (mode, hist, w, xh)."""

    def __init__(self, fps, n_min=0, n_max=30, n_f=3):
        self.n_f = n_f
        self.n_i = horizon_sizes(n_min, n_max, n_f)
        self.gains = [lsf_gain(n, 1 / fps) for n in self.n_i]

    def _estimates(self, st):
        # gsff.py:230-240: x_hat[:, i] = (gain_i @ flatten(last n_i measurements))[:2]
        for i in range(st['mode']):
            y = [v for m in st['hist'][-self.n_i[i]:] for v in m]
            st['xh'][:, i] = np.dot(self.gains[i], y)[:2]

    def correct(self, st, z):
        """gsff.py:251-347.  ``st`` is {} on a track's first call."""
        if 'hist' not in st:
            st.update(mode=0, hist=[z] * self.n_i[0], w=None, xh=None)      # gsff.py:279-281
        switched = False
        if st['mode'] < self.n_f:                                           # gsff.py:284-289
            while len(st['hist']) >= self.n_i[st['mode']]:
                st['mode'] += 1
                switched = True
                if st['mode'] >= self.n_f:
                    break
        if switched:                                                        # gsff.py:291-308
            m = st['mode']
            st['xh'] = np.zeros((2, m))
            st['w'] = 1 / m * np.ones(m)
            self._estimates(st)
        m = st['mode']
        lik = []
        for i in range(m):                                                  # gsff.py:179-202, 310-313
            d = z - st['xh'][:, i]
            v = np.exp(-0.5 * np.dot(d.T, np.dot(np.eye(2), d)))
            lik.append(LIKELIHOOD_MIN if v < LIKELIHOOD_MIN else v)
        st['hist'].append(z)                                                # gsff.py:315-318
        if len(st['hist']) > self.n_i[-1] + 1:
            st['hist'] = st['hist'][-(self.n_i[-1] + 1):]
        total = sum(lik * st['w'])                                          # gsff.py:321
        for i in range(m):                                                  # gsff.py:332-334
            st['w'][i] = lik[i] * st['w'][i] / total
        return np.sum(st['xh'] * st['w'], axis=1)                           # gsff.py:337

    def predict(self, st):
        """gsff.py:204-249."""
        self._estimates(st)
        return np.sum(st['xh'] * st['w'], axis=1)


class LinkerPort:
    """Array-based restatement of CentroidTracker (tracker.py:27-230).

    Tracks are kept in insertion order in parallel lists (the reference uses OrderedDicts keyed by id; deleting
    a key keeps the order of the rest, so a list with in-place removal is equivalent)."""

    def __init__(self, max_disappeared, fps=30.0, n_min=0, n_max=30, n_f=3, use_gsff=True):
        self.max_disappeared = max_disappeared
        self.use_gsff = use_gsff
        self.next_id = 0
        self.ids = []       # track ids, insertion order
        self.pos = []       # np.float64[2]: measurement / GSFF prediction used for the next association
        self.info = []      # (w, h, deg) or [0, 0, 0]
        self.gone = []      # consecutive-miss counters
        self.filt = []      # per-track GSFF state dicts
        if use_gsff:
            self.bank = GsffBank(fps, n_min, fps if n_max is None else n_max, n_f)

    # -- lifecycle ------------------------------------------------------------------------------------
    def _register(self, xy, info):                                          # tracker.py:73-82
        self.ids.append(self.next_id); self.pos.append(xy); self.info.append(info)
        self.gone.append(0); self.filt.append({})
        self.next_id += 1

    def _age(self, rows):                                                   # tracker.py:99-107, 200-211
        drop = []
        for r in rows:
            self.gone[r] += 1
            self.info[r] = [0] * len(self.info[r])
            if self.gone[r] > self.max_disappeared:
                drop.append(r)
        for r in sorted(drop, reverse=True):
            for lst in (self.ids, self.pos, self.info, self.gone, self.filt):
                del lst[r]

    # -- one frame --------------------------------------------------------------------------------------
    def update(self, rects):
        """rects = [((x, y), (w, h, deg)), ...] in contour order.  Returns [(id, xy_out, info)] for every live
        track in insertion order (what track_eval.py:313-316 appends to ``coords``)."""
        if len(rects) == 0:
            self._age(list(range(len(self.ids))))
        else:
            dets = np.zeros((len(rects), 2), dtype='float')
            for i, (xy, _) in enumerate(rects):
                dets[i] = xy
            if not self.ids:
                for i in range(len(rects)):
                    self._register(dets[i], rects[i][1])
            else:
                dm = cdist(np.array(self.pos), dets)                        # tracker.py:151
                rows = dm.min(axis=1).argsort()                             # tracker.py:158
                cols = dm.argmin(axis=1)[rows]                              # tracker.py:163
                taken_r, taken_c = set(), set()
                for r, c in zip(rows, cols):                                # tracker.py:171-189
                    if r in taken_r or c in taken_c:
                        continue
                    self.pos[r] = dets[c]; self.info[r] = rects[c][1]; self.gone[r] = 0
                    taken_r.add(r); taken_c.add(c)
                n, m = dm.shape
                if n >= m:                                                  # tracker.py:198-211
                    self._age([r for r in set(range(n)).difference(taken_r)])
                else:                                                       # tracker.py:215-217
                    for c in unused_cols_order(m, taken_c):
                        self._register(dets[c], rects[c][1])
        out = []
        for r in range(len(self.ids)):                                      # tracker.py:219-230
            if self.use_gsff:
                xy = self.bank.correct(self.filt[r], self.pos[r])
                self.pos[r] = self.bank.predict(self.filt[r])
            else:
                xy = self.pos[r]
            out.append((self.ids[r], xy, self.info[r]))
        return out
