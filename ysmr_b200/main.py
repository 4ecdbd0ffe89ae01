"""``ysmr()`` / ``analyse()`` with the reference's signatures (ysmr/main.py:32-172, 175-331), routing videos to the B200
``track_bacteria``.

Of the stages after tracking, ``select_tracks`` runs on the GPU as well (``ysmr_b200/select.py``);
``evaluate_tracks`` and ``annotate_video`` are out of scope of this repository and are NOT
re-implemented: when the reference package ``ysmr`` is importable, :func:`install` rebinds its ``track_bacteria`` and
``select_tracks`` to ours and these wrappers simply call the reference's own drivers, so every later stage, log line and
file of the reference is produced by the reference's code on our ``_list.csv`` / selected rows.  Without the reference
package the wrappers run tracking and selection and keep the return conventions (DataFrame / True / None; list of
``(path, result)``).
"""
from __future__ import annotations

import logging
import os

from .select import select_tracks
from .settings import get_configs
from .track_eval import track_bacteria

__all__ = ['ysmr', 'analyse', 'install', 'uninstall']

_saved = {}


def install():
    """Make the reference use the GPU path: rebinds ``track_bacteria`` and ``select_tracks`` in ``ysmr.main`` (imported by
    name at main.py:27) and in ``ysmr.track_eval``.  Returns True when the reference package was found."""
    try:
        import ysmr.main as ref_main
        import ysmr.track_eval as ref_te
    except Exception:
        return False
    if 'main' not in _saved:
        _saved['main'] = ref_main.track_bacteria
        _saved['te'] = ref_te.track_bacteria
        _saved['main_sel'] = ref_main.select_tracks
        _saved['te_sel'] = ref_te.select_tracks
    ref_main.track_bacteria = track_bacteria
    ref_te.track_bacteria = track_bacteria
    ref_main.select_tracks = select_tracks
    ref_te.select_tracks = select_tracks
    return True


def uninstall():
    if 'main' in _saved:
        import ysmr.main as ref_main
        import ysmr.track_eval as ref_te
        ref_main.track_bacteria = _saved.pop('main')
        ref_te.track_bacteria = _saved.pop('te')
        ref_main.select_tracks = _saved.pop('main_sel')
        ref_te.select_tracks = _saved.pop('te_sel')


def analyse(path, settings=None, result_folder=None, return_df=False, **kwargs):
    if install():
        import ysmr.main as ref_main
        return ref_main.analyse(path, settings=settings, result_folder=result_folder, return_df=return_df, **kwargs)
    logger = logging.getLogger('ysmr').getChild(__name__)
    settings = get_configs(settings)
    if settings is None:
        return None
    if result_folder is None:
        result_folder = os.path.dirname(os.path.abspath(path))
    if any(ext in path for ext in ('_analysed.csv', '_statistics.csv', '_annotated_output.')):
        logger.warning('File already evaluated. File: {}'.format(path))
        return None
    # main.py:86-127: a video is tracked first, a *_list.csv goes straight to the selection, a *_selected_data.csv is done
    df, csv_file, meta = None, None, dict(kwargs)
    if '.csv' not in path:
        res = track_bacteria(video_path=path, settings=settings, result_folder=result_folder)
        if res is None:
            logger.warning('Error during video analysis of file {}.'.format(path))
            return None
        df, fps, height, width, csv_file = res
        meta.update(fps=fps, frame_height=height, frame_width=width)
        path = csv_file
    if 'selected_data.csv' not in path:
        df = select_tracks(path_to_file=path, df=df, results_directory=result_folder, settings=settings, **meta)
        if df is None:
            logger.warning('Error during video analysis of file {}.'.format(path))
    else:
        # an already selected file: the reference would go on to evaluate_tracks (main.py:130-138), which this repository
        # does not replace -- the rows are handed back as they are
        import pandas as pd
        try:
            df = pd.read_csv(path)
        except Exception as ex:
            logger.critical('Error reading data frame from file {}: {!r}'.format(path, ex))
            df = None
    if settings.get('delete .csv file after analysis') and csv_file:
        try:
            os.remove(csv_file)
        except OSError:
            pass
    if df is None:
        return None
    return df if return_df else True


def ysmr(paths=None, settings=None, result_folder=None, multiprocess=False):
    if install():
        import ysmr.main as ref_main
        # process-per-video (main.py:281-288) would need the rebinding in every child; videos are GPU-bound anyway
        if multiprocess:
            logging.getLogger('ysmr').getChild(__name__).warning(
                'multiprocess=True is ignored by the B200 path: videos are processed one after the other on the GPU.')
        return ref_main.ysmr(paths=paths, settings=settings, result_folder=result_folder, multiprocess=False)
    settings = get_configs(settings)
    if settings is None:
        print('Fatal error in retrieving tracking.ini')
        return None
    if isinstance(paths, (str, os.PathLike)):
        paths = [paths]
    if not paths:
        return None
    if multiprocess:
        logging.getLogger('ysmr').getChild(__name__).warning(
            'multiprocess=True is ignored by the B200 path: videos are processed one after the other on the GPU.')
    finished = []
    for p in [os.path.expanduser(q) for q in paths]:
        finished.append((p, analyse(p, settings=settings, result_folder=result_folder)))
    return finished
