"""``track_bacteria`` -- drop-in for the reference's function of the same name (ysmr/track_eval.py:38-405).

Same signature, same return tuple ``(df_sorted, fps_of_file, frame_height, frame_width, list_csv_path)`` or ``None`` on
any failure, same ``<result_folder>/<stem>_list.csv``.  What changes is the body of the ``while True`` loop
(track_eval.py:156-366): frames are decoded by cv2.VideoCapture exactly as before, but by a reader thread into one of two
pinned host buffers, and handed, a chunk at a time, to ``ysmr_track_host`` (include/ysmr_b200.h), which runs detection
and linking on the GPU (H2D copies double-buffered under the kernels) and returns the rows the loop would have appended
to ``coords``; decoding of the next chunk overlaps the GPU work of the current one.

Unsupported on this path (the function logs and returns None rather than silently diverging):
'include luminosity in tracking calculation' (broken with GSFF in the reference itself), colour
filters other than COLOR_BGR2GRAY, and the interactive display ('display video analysis' is ignored with a warning --
there is nothing per-frame on the host to draw on).
"""
from __future__ import annotations

import logging
import os

import numpy as np

from . import listio
from .settings import COLOR_BGR2GRAY, get_configs

__all__ = ['track_bacteria']


def create_results_folder(path):
    """Dated result folder next to the video, like helper_file.create_results_folder (377-405)."""
    from time import localtime, strftime
    directory = os.path.abspath(os.path.join(os.path.dirname(path), '{}_Results/'.format(strftime('%y%m%d', localtime()))))
    try:
        os.makedirs(directory, exist_ok=True)
    except OSError:
        directory = './'
    return directory


def _pinned_frames(n, h, w):
    import torch
    t = torch.empty((n, h, w, 3), dtype=torch.uint8)
    try:
        t = t.pin_memory()
    except Exception:                     # pinning is an optimisation only
        pass
    return t, t.numpy()


def track_bacteria(video_path, settings=None, result_folder=None, *, device=0, chunk_frames=256, max_blobs=4096,
                   max_tracks=8192, row_sink='append'):
    import cv2
    from .api import ROW_DTYPE, Context
    logger = logging.getLogger('ysmr').getChild(__name__)
    settings = get_configs(settings)
    if settings is None:
        logger.critical('No settings provided / could not get settings for start_it_up().')
        return None
    if not os.path.isfile(video_path):
        logger.critical('File {} does not exist'.format(video_path))
        return None
    if row_sink not in ('append', 'once'):
        logger.critical("row_sink must be 'append' or 'once', not {!r}".format(row_sink))
        return None
    try:
        cap = cv2.VideoCapture(video_path)
    except (IOError, OSError, cv2.error) as err:
        logger.exception('Cannot open file {} due to error: {}'.format(video_path, err))
        return None
    frame_count = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    if frame_count < settings['minimal frame count']:
        logger.warning('File {} too short; file was skipped. Limit for \'minimal frame count\': {}'.format(
            video_path, settings['minimal frame count']))
        return None
    if not settings['force tracking.ini fps settings']:
        fps_of_file = cap.get(cv2.CAP_PROP_FPS)
        if settings['verbose'] or fps_of_file != settings['frames per second']:
            logger.info('fps of file: {}'.format(fps_of_file))
    else:
        fps_of_file = settings['frames per second']
    if not fps_of_file or fps_of_file <= 0:
        logger.critical('fps unacceptable: {}'.format(fps_of_file))
        return None
    if settings['include luminosity in tracking calculation']:
        logger.critical('\'include luminosity in tracking calculation\' is not supported by the B200 path.')
        return None
    if settings['color filter'] != COLOR_BGR2GRAY:
        logger.critical('Only COLOR_BGR2GRAY is supported by the B200 path.')
        return None
    if settings['display video analysis']:
        logger.warning('\'display video analysis\' is ignored by the B200 path.')
    if not result_folder:
        result_folder = create_results_folder(video_path)
    logger.info('Starting with file {}'.format(video_path))
    old_list, list_name = listio.start_list(video_path, result_folder, settings['rename previous result .csv'])

    # track_eval.py:127-132 -- including the in-place mutation of the caller's dict
    white = bool(settings['white bacteria on dark background'])
    if not white:
        settings['threshold offset for detection'] = settings['threshold offset for detection'] * -1
    signed_offset = settings['threshold offset for detection']
    frame_height, frame_width = int(cap.get(4)), int(cap.get(3))

    try:
        ctx = Context(frame_height, frame_width, 3, device,
                      white_on_dark=white, offset=signed_offset if white else -signed_offset,
                      adt=settings['adaptive double threshold'], fps=fps_of_file,
                      use_gsff=not settings['disable gsff'], n_f=settings['number of LSFFs'],
                      n_min=settings['minimum horizon size'], n_max=settings['maximum horizon size'],
                      max_batch=min(chunk_frames, 256), max_blobs=max_blobs, max_tracks=max_tracks)
    except Exception as ex:
        logger.critical('Could not create the GPU context: {}'.format(ex))
        cap.release()
        return None

    # Ingest: a reader thread decodes chunk i+1 into the other pinned buffer while the GPU works on chunk i
    # (cv2.VideoCapture.read and the ctypes call both release the GIL); the queues hand the two buffers back and forth.
    import queue
    import threading
    keep = [_pinned_frames(chunk_frames, frame_height, frame_width) for _ in range(2)]
    free_q, full_q = queue.Queue(), queue.Queue()
    for i in range(2):
        free_q.put(i)
    stop_reading = threading.Event()

    def reader():
        try:
            while not stop_reading.is_set():
                i = free_q.get()
                if i is None:
                    return
                b = keep[i][1]
                n = 0
                while n < chunk_frames and not stop_reading.is_set():
                    ret, frame = cap.read()
                    if not ret:
                        full_q.put((i, n, True))
                        return
                    b[n] = frame
                    n += 1
                full_q.put((i, n, False))
        except Exception as ex:                      # surfaces in the consumer
            full_q.put(ex)

    rows_out = np.empty(min(chunk_frames * max_tracks, 1 << 22), ROW_DTYPE)
    pending = []            # rows not yet written (flushed every 'list save length interval' rows like the reference)
    n_pending = 0
    curr_frame_count = 0
    error_during_read = False
    last_live = 0
    done = False
    th = threading.Thread(target=reader, name='ysmr-b200-decode', daemon=True)
    th.start()
    try:
        while not done:
            item = full_q.get()
            if isinstance(item, Exception):
                raise item
            i, n, done = item
            buf = keep[i][1]
            if done:
                total = curr_frame_count + n
                if (frame_count == total + 1 or frame_count == total) and frame_count >= settings['minimal frame count']:
                    logger.debug('Frames from file {} read.'.format(os.path.basename(video_path)))
                else:
                    logger.critical('Error during cap.read() with file {}'.format(video_path))
                    error_during_read = settings['stop evaluation on error']
            if n:
                rows = ctx.track_host(buf[:n], curr_frame_count, rows_capacity=len(rows_out), rows_out=rows_out)
                free_q.put(i)
                pending.append(rows.copy()); n_pending += len(rows)
                curr_frame_count += n
                last_live = ctx.live_tracks()[0]
                # row_sink 'append' (default): the reference's life cycle, text appended every 'list save length interval'
                # rows and sorted through a read-back at the end; 'once': rows stay in memory and the sorted file is
                # written a single time (listio.write_sorted) -- same bytes, no intermediate file traffic
                if row_sink == 'append' and n_pending >= settings['list save length interval']:
                    listio.append_rows(list_name, np.concatenate(pending))
                    pending, n_pending = [], 0
        if pending and row_sink == 'append':
            listio.append_rows(list_name, np.concatenate(pending))
    except Exception as ex:
        logger.exception('GPU tracking failed for {}: {}'.format(video_path, ex))
        stop_reading.set(); free_q.put(None); th.join(timeout=30)
        cap.release()
        ctx.close()
        return None
    stop_reading.set(); free_q.put(None); th.join(timeout=30)
    cap.release()
    ctx.close()
    del keep

    if old_list and error_during_read:
        try:
            os.remove(list_name)
            os.rename(old_list, list_name)
            logger.info('Restoring old list: {}'.format(list_name))
        except OSError as err:
            logger.error('Could not restore {}: {!r}'.format(list_name, err.args))
    if last_live == 0:            # track_eval.py:388-392: no object alive after the last frame
        logger.warning('Did not track any objects. File: {}'.format(video_path))
        return None
    if row_sink == 'append':
        df_for_eval = listio.sort_list(list_name, save_file=not settings['delete .csv file after analysis'])
    else:
        all_rows = np.concatenate(pending) if pending else np.empty(0, ROW_DTYPE)
        df_for_eval = listio.write_sorted(list_name, all_rows, save_file=not settings['delete .csv file after analysis'])
    logger.info('frames: {:>6} of {:>6}, csv: {}'.format(curr_frame_count, frame_count, list_name))
    if error_during_read:
        logger.critical('Error during read, stopping before evaluation. File: {}'.format(video_path))
        return None
    return df_for_eval, fps_of_file, frame_height, frame_width, list_name
