"""Ingest for the drop-in: video file -> pinned host buffers, decoded ahead of the GPU.

The reference reads frames one at a time on the thread that also runs the whole pipeline (``cap.read()``,
track_eval.py:65,159).  With detection and linking on the GPU the wall time of ``track_bacteria`` IS the decode time, so:

* frames are decoded by reader threads into pinned buffers while the GPU works on the previous chunk (cv2.VideoCapture.read
  and the ctypes call into the library both release the GIL);
* for intra-only codecs (every frame a key frame: FFV1, HuffYUV, raw Y800, PNG, MJPEG -- what lossless microscope
  recordings use) several readers decode disjoint chunks in parallel, each with its own ``cv2.VideoCapture`` seeked to the
  chunk start; other codecs get the single sequential reader (a seek there may land on a different frame);
* grey sources (B == G == R in every decoded frame, what ``cap.read()`` delivers for grey codecs) are handed over as ONE
  plane: BGR2GRAY is the identity on such pixels, so one third of the bytes crosses PCIe.  Every frame is
  checked; the first colour frame makes the source fall back to three planes for the whole video.

Chunks are delivered strictly in frame order whatever order they were decoded in.
"""
from __future__ import annotations

import queue
import threading

import numpy as np

# compared case-insensitively (FFmpeg reports 'ffv1' for a file written as 'FFV1')
INTRA_ONLY_FOURCC = {'ffv1', 'hfyu', 'ffvh', 'y800', 'grey', 'y8  ', 'png ', 'mpng', 'mjpg', 'jpeg', 'i420', 'iyuv', 'yv12',
                     'dib ', 'rgba', 'bgr3', 'rgb3', '\x00\x00\x00\x00'}


def fourcc_of(cap):
    import cv2
    v = int(cap.get(cv2.CAP_PROP_FOURCC))
    return ''.join(chr((v >> (8 * i)) & 0xFF) for i in range(4))


def is_grey_frame(frame):
    """B == G == R for every pixel of a decoded BGR frame."""
    return bool(np.array_equal(frame[..., 0], frame[..., 1]) and np.array_equal(frame[..., 1], frame[..., 2]))


class ColourFrame(Exception):
    """A source that started out grey delivered a colour frame (the caller restarts with three planes)."""


def pinned(shape):
    import torch
    t = torch.empty(shape, dtype=torch.uint8)
    try:
        t = t.pin_memory()
    except Exception:                     # pinning is an optimisation only
        pass
    return t, t.numpy()


class ChunkReader:
    """Iterates ``(buffer_view, n_frames, is_last)`` over a video in frame order; ``release(view)`` hands a buffer back.

    channels = 3: frames as ``cap.read()`` delivers them; channels = 1: plane 0 of frames verified to be grey.
    """

    def __init__(self, video_path, frame_count, height, width, channels, chunk_frames, n_readers=1, n_buffers=None):
        import cv2
        self.path, self.frame_count = video_path, int(frame_count)
        self.h, self.w, self.channels, self.chunk = int(height), int(width), int(channels), int(chunk_frames)
        cap = cv2.VideoCapture(video_path)
        intra = fourcc_of(cap).lower() in INTRA_ONLY_FOURCC
        cap.release()
        self.n_readers = max(1, int(n_readers)) if intra and self.frame_count > 0 else 1
        n_buffers = n_buffers or (self.n_readers + 1)
        shape = (self.chunk, self.h, self.w) if self.channels == 1 else (self.chunk, self.h, self.w, 3)
        self._keep = [pinned(shape) for _ in range(max(2, n_buffers))]
        self._free = queue.Queue()
        for i in range(len(self._keep)):
            self._free.put(i)
        self._done = {}                    # chunk index -> (buffer index, n, exhausted) or an Exception
        self._cv = threading.Condition()
        self._stop = threading.Event()
        self._next_chunk = 0
        self._lock = threading.Lock()
        n_chunks = (self.frame_count + self.chunk - 1) // self.chunk if self.n_readers > 1 else None
        self._n_chunks = n_chunks
        self._threads = [threading.Thread(target=self._run, args=(r,), name=f'ysmr-b200-decode-{r}', daemon=True)
                         for r in range(self.n_readers)]
        for t in self._threads:
            t.start()

    # -- reader threads -----------------------------------------------------------------------------------------------
    def _store(self, i, n, frame):
        b = self._keep[i][1]
        if self.channels == 1:
            if not is_grey_frame(frame):
                raise ColourFrame()
            b[n] = frame[..., 0]
        else:
            b[n] = frame

    def _publish(self, c, item):
        with self._cv:
            self._done[c] = item
            self._cv.notify_all()

    def _run(self, r):
        import cv2
        cap = cv2.VideoCapture(self.path)
        pos = 0                            # frame the capture will deliver next
        c = -1
        try:
            while not self._stop.is_set():
                # buffer first, chunk index second: every claimed chunk owns a buffer, so the chunk the consumer waits for
                # can always be decoded (no deadlock with later chunks holding all buffers)
                i = self._free.get()
                if i is None:
                    return
                with self._lock:
                    c = self._next_chunk
                    if self._n_chunks is not None and c >= self._n_chunks:
                        self._free.put(i)
                        return
                    self._next_chunk += 1
                start = c * self.chunk
                if self.n_readers > 1 and pos != start:
                    cap.set(cv2.CAP_PROP_POS_FRAMES, start)
                    pos = start
                n, exhausted = 0, False
                while n < self.chunk and not self._stop.is_set():
                    ret, frame = cap.read()
                    if not ret:
                        exhausted = True
                        break
                    self._store(i, n, frame)
                    n += 1
                    pos += 1
                self._publish(c, (i, n, exhausted))
                if exhausted:
                    return
        except Exception as ex:            # surfaces in the consumer, in frame order
            self._publish(max(c, 0), ex)
        finally:
            cap.release()

    # -- consumer -----------------------------------------------------------------------------------------------------
    def __iter__(self):
        c = 0
        while True:
            with self._cv:
                while c not in self._done:
                    if self._n_chunks is not None and c >= self._n_chunks:
                        return
                    self._cv.wait(timeout=0.5)
                    if self._stop.is_set():
                        return
                item = self._done.pop(c)
            if isinstance(item, Exception):
                raise item
            i, n, exhausted = item
            last = exhausted or (self._n_chunks is not None and c + 1 >= self._n_chunks)
            yield self._keep[i][1][:n], i, n, last
            if last:
                return
            c += 1

    def release(self, i):
        self._free.put(i)

    def close(self):
        self._stop.set()
        for _ in self._threads:
            self._free.put(None)
        with self._cv:
            self._cv.notify_all()
        for t in self._threads:
            t.join(timeout=30)
        self._keep = []
