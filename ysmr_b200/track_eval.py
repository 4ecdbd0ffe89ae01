"""``track_bacteria`` -- drop-in for the reference's function of the same name (ysmr/track_eval.py:38-405).

Same signature, same return tuple ``(df_sorted, fps_of_file, frame_height, frame_width, list_csv_path)`` or ``None`` on
any failure, same ``<result_folder>/<stem>_list.csv``.  What changes is the body of the ``while True`` loop
(track_eval.py:156-366): frames are decoded by cv2.VideoCapture exactly as before, but by reader threads into pinned host
buffers (``ysmr_b200/ingest.py``: several decoders in parallel for intra-only codecs, one plane instead of three for grey
sources), and handed, a chunk at a time, to ``ysmr_track_host`` (include/ysmr_b200.h), which runs detection and linking on
the GPU (H2D copies double-buffered under the kernels) and returns the rows the loop would have appended to ``coords``;
decoding of the coming chunks overlaps the GPU work of the current one.

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


def track_bacteria(video_path, settings=None, result_folder=None, *, device=0, chunk_frames=None, max_blobs=4096,
                   max_tracks=8192, row_sink='append', n_readers=None, grey_source='auto'):
    """Extra keyword arguments (all optional, none changes the result): device; chunk_frames (frames per call into the
    library, default: about 256 MB of frames); max_blobs / max_tracks (device capacities; exceeding one fails loudly);
    row_sink ('append' = the reference's file life cycle, 'once' = rows sorted on the device and written a single time);
    n_readers (decoder threads for intra-only codecs, default min(4, cpu count)); grey_source ('auto': one plane is uploaded
    while every decoded frame has B == G == R, three planes otherwise; 'bgr': always three)."""
    import cv2
    from . import ingest
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
    try:
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
        frame_height, frame_width = int(cap.get(4)), int(cap.get(3))
        first_is_grey = False
        if grey_source == 'auto':
            ret, frame0 = cap.read()
            first_is_grey = bool(ret) and frame0.ndim == 3 and ingest.is_grey_frame(frame0)
    finally:
        cap.release()
    if not result_folder:
        result_folder = create_results_folder(video_path)
    logger.info('Starting with file {}'.format(video_path))
    old_list, list_name = listio.start_list(video_path, result_folder, settings['rename previous result .csv'])

    # track_eval.py:127-132 -- including the in-place mutation of the caller's dict
    white = bool(settings['white bacteria on dark background'])
    if not white:
        settings['threshold offset for detection'] = settings['threshold offset for detection'] * -1
    signed_offset = settings['threshold offset for detection']
    if chunk_frames is None:                     # about 256 MB of BGR frames per buffer, whatever the resolution
        chunk_frames = int(max(16, min(256, (256 << 20) // max(1, frame_height * frame_width * 3))))
    if n_readers is None:
        n_readers = min(4, os.cpu_count() or 1)

    def restore_old_list(always_remove_partial):
        """track_eval.py:378-387 (old list back on a read error); after a failure of the GPU path the partial list is
        removed as well, so that a later analyse() over the folder cannot pick up a truncated csv."""
        try:
            if os.path.isfile(list_name) and (old_list or always_remove_partial):
                os.remove(list_name)
            if old_list:
                os.rename(old_list, list_name)
                logger.info('Restoring old list: {}'.format(list_name))
        except OSError as err:
            logger.error('Could not restore {}: {!r}'.format(list_name, err.args))

    def run(channels):
        """One pass over the video; returns (rows or None, frames read, live tracks, read error)."""
        ctx = reader = None
        try:
            ctx = Context(frame_height, frame_width, channels, device,
                          white_on_dark=white, offset=signed_offset if white else -signed_offset,
                          adt=settings['adaptive double threshold'], fps=fps_of_file,
                          use_gsff=not settings['disable gsff'], n_f=settings['number of LSFFs'],
                          n_min=settings['minimum horizon size'], n_max=settings['maximum horizon size'],
                          max_batch=min(chunk_frames, 256), max_blobs=max_blobs, max_tracks=max_tracks)
            # Ingest: reader threads decode the coming chunks into pinned buffers while the GPU works on the
            # current one (cv2.VideoCapture.read and the ctypes call both release the GIL)
            reader = ingest.ChunkReader(video_path, frame_count, frame_height, frame_width, channels, chunk_frames, n_readers)
            rows_cap = min(chunk_frames * max_tracks, 1 << 22)
            once = row_sink == 'once'
            rows_out = None if once else np.empty(rows_cap, ROW_DTYPE)
            if once:                       # rows stay on the device and come back sorted, once
                ctx.archive_rows(True)
            pending, n_pending = [], 0     # rows not yet written (flushed every 'list save length interval' rows)
            curr, last_live, read_error = 0, 0, False
            for buf, idx, n, last in reader:
                if last:
                    total = curr + n
                    if (frame_count == total + 1 or frame_count == total) and frame_count >= settings['minimal frame count']:
                        logger.debug('Frames from file {} read.'.format(os.path.basename(video_path)))
                    else:
                        logger.critical('Error during cap.read() with file {}'.format(video_path))
                        read_error = settings['stop evaluation on error']
                if n:
                    if once:
                        ctx.track_host(buf, curr, rows_capacity=rows_cap, copy_rows=False)
                    else:
                        rows = ctx.track_host(buf, curr, rows_capacity=rows_cap, rows_out=rows_out)
                        pending.append(rows.copy()); n_pending += len(rows)
                    curr += n
                    last_live = ctx.live_tracks()[0]
                    # row_sink 'append' (default): the reference's life cycle, text appended every 'list save length
                    # interval' rows and sorted through a read-back at the end; 'once': rows stay on the device, are
                    # grouped by (TRACK_ID, POSITION_T) there and written a single time -- same bytes
                    if row_sink == 'append' and n_pending >= settings['list save length interval']:
                        listio.append_rows(list_name, np.concatenate(pending))
                        pending, n_pending = [], 0
                reader.release(idx)
            if pending and row_sink == 'append':
                listio.append_rows(list_name, np.concatenate(pending))
                pending = []
            return (ctx.rows_sorted() if once else None), curr, last_live, read_error
        finally:
            if reader is not None:
                reader.close()
            if ctx is not None:
                ctx.close()

    try:
        try:
            pending, curr_frame_count, last_live, error_during_read = run(1 if first_is_grey else 3)
        except ingest.ColourFrame:
            logger.warning('File {} starts grey but contains colour frames: restarting with three planes.'.format(video_path))
            listio.reset_list(list_name)
            pending, curr_frame_count, last_live, error_during_read = run(3)
    except Exception as ex:
        # device capacities (max_blobs, max_tracks, runs, rows) are reported with the first frame that exceeded them
        logger.exception('GPU tracking failed for {}: {}'.format(video_path, ex))
        restore_old_list(True)
        return None

    if error_during_read:
        restore_old_list(False)
    if last_live == 0:            # track_eval.py:388-392: no object alive after the last frame
        logger.warning('Did not track any objects. File: {}'.format(video_path))
        return None
    if error_during_read:
        logger.critical('Error during read, stopping before evaluation. File: {}'.format(video_path))
        return None
    if row_sink == 'append':
        df_for_eval = listio.sort_list(list_name, save_file=not settings['delete .csv file after analysis'])
    else:
        df_for_eval = listio.write_sorted(list_name, pending, save_file=not settings['delete .csv file after analysis'],
                                          presorted=True)
    logger.info('frames: {:>6} of {:>6}, csv: {}'.format(curr_frame_count, frame_count, list_name))
    return df_for_eval, fps_of_file, frame_height, frame_width, list_name
