"""``select_tracks`` with the reference's signature and return conventions (ysmr/track_eval.py:541-843), the data-parallel
body on the GPU (``ysmr_select_tracks``, csrc/select.cu / select.cuh).

What stays on the host is what the reference does once per file: the prologue checks and their log lines (:558-606), the
log lines that report the quantiles and the kick reasons (:686-691, 709, 723-741, 796-812), building the returned
DataFrame (:822-835) and writing ``<name>_selected_data.csv`` (:836-838).  There is no CPU fallback: without the CUDA
library or a GPU the call raises.
"""
from __future__ import annotations

import ctypes as C
import logging
import os

import numpy as np

from . import _lib
from .listio import DTYPES

# the keys select_tracks reads, with the reference's defaults AFTER get_configs (helper_file.py:160-282, 586-843)
SELECT_DEFAULTS = {
    'verbose': False,
    'frames per second': 30.0,
    'frame height': 922,
    'frame width': 1228,
    'pixel per micrometre': 1.41888781,
    'force tracking.ini fps settings': False,
    'minimal length in seconds': 20.0,
    'limit track length to x seconds': 20.0,
    'limit track length exactly': False,
    'extreme area outliers lower end in px*px': 2,
    'extreme area outliers upper end in px*px': 50,
    'exclude measurement when above x times average area': 1.5,
    'maximal consecutive holes': 5,
    'maximal empty frames in %': 5.0 / 100 + 1,
    'percent quantiles excluded area': 10.0 / 100,
    'try to omit motility outliers': True,
    'stop excluding motility outliers if total count above percent': 5.0 / 100,
    'average width/height ratio min.': 0.125,          # 'rod shaped bacteria' preset (helper_file.py:634-639)
    'average width/height ratio max.': 0.67,
    'percent of screen edges to exclude': 5.0 / 100,
    'maximal recursion depth': 960,
    'store processed .csv file': True,
    'path to test .csv': '',
}

COLUMNS = ['TRACK_ID', 'POSITION_T', 'POSITION_X', 'POSITION_Y', 'WIDTH', 'HEIGHT', 'DEGREES_ANGLE']


def _settings(settings):
    """dict -> the same dict with missing selection keys filled in; ini path / None -> the reference's get_configs when the
    reference package is importable (it knows every key), else the hot-path parser plus the defaults above."""
    if isinstance(settings, dict):
        for key, val in SELECT_DEFAULTS.items():
            settings.setdefault(key, val)
        return settings
    from .settings import get_configs
    out = get_configs(settings)
    if out is None:
        return None
    for key, val in SELECT_DEFAULTS.items():
        out.setdefault(key, val)
    return out


def select_params(settings, fps, frame_height, frame_width):
    """ysmr_select_params from the settings dict (the conversions of track_eval.py:586-587)."""
    p = _lib.SelectParams()
    p.area_lo = float(settings['extreme area outliers lower end in px*px'])
    p.area_hi = float(settings['extreme area outliers upper end in px*px'])
    p.area_factor = float(settings['exclude measurement when above x times average area'] or 0.0)
    p.q_area = float(settings['percent quantiles excluded area'])
    p.stop_outliers_above = float(settings['stop excluding motility outliers if total count above percent'])
    p.max_empty = float(settings['maximal empty frames in %'])
    p.ratio_min = float(settings['average width/height ratio min.'])
    p.ratio_max = float(settings['average width/height ratio max.'])
    p.edge = float(settings['percent of screen edges to exclude'])
    p.min_len_frames = int(round(fps, 0) * settings['minimal length in seconds'])
    p.limit_frames = int(round(fps, 0) * settings['limit track length to x seconds'])
    p.limit_exactly = 1 if settings['limit track length exactly'] else 0
    p.omit_motility_outliers = 1 if settings['try to omit motility outliers'] else 0
    p.max_holes = int(settings['maximal consecutive holes'])
    p.max_recursion = int(settings['maximal recursion depth'])
    p.frame_h = int(frame_height)
    p.frame_w = int(frame_width)
    return p


def select_rows(track_id, t, x, y, w, h, params, device=0):
    """The C-ABI call on plain arrays.  Returns (good[u8], clean_index[i32], kick_reasons[9], info[16])."""
    lib = _lib.load()
    n = len(track_id)
    cols = [np.ascontiguousarray(track_id, np.uint32), np.ascontiguousarray(t, np.uint32)] + \
           [np.ascontiguousarray(a, np.float64) for a in (x, y, w, h)]
    good = np.zeros(n, np.uint8)
    clean_index = np.empty(n, np.int32)
    kicks = np.zeros(9, np.int64)
    info = np.zeros(_lib.SELECT_INFO, np.float64)
    rc = lib.ysmr_select_tracks(int(device), n, *[a.ctypes.data_as(C.c_void_p) for a in cols], C.byref(params),
                                good.ctypes.data_as(C.c_void_p), clean_index.ctypes.data_as(C.c_void_p),
                                kicks.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError('ysmr_select_tracks failed ({}): {}'.format(rc, lib.ysmr_select_last_error().decode()))
    return good, clean_index, kicks, info


def select_tracks(path_to_file=None, df=None, results_directory=None, fps=None, frame_height=None, frame_width=None,
                  settings=None, device=0, **_):
    """Drop-in for ysmr.track_eval.select_tracks: returns the DataFrame of the selected rows ('index' + the seven columns)
    or None, and writes <name>_selected_data.csv when 'store processed .csv file' is set."""
    import pandas as pd
    logger = logging.getLogger('ysmr').getChild(__name__)
    settings = _settings(settings)
    if settings is None:
        logger.critical('No settings provided / could not get settings for start_it_up().')
        return None
    if path_to_file is None:
        path_to_file = settings['path to test .csv']
    if results_directory is None:
        results_directory = os.path.dirname(os.path.abspath(path_to_file))
    file_name = os.path.splitext(os.path.basename(path_to_file))[0]
    if fps is None or fps <= 0 or settings['force tracking.ini fps settings']:
        if settings['frames per second'] > 0:
            fps = settings['frames per second']
        else:
            logger.critical('fps value is negative or zero; cannot continue.')
            return None
    if settings['extreme area outliers lower end in px*px'] >= settings['extreme area outliers upper end in px*px']:
        logger.critical('Minimal area exclusion in px^2 larger or equal to maximum; will not be able to find tracks. '
                        'Please update tracking.ini. extreme area outliers lower end in px*px: {}, '
                        'extreme area outliers upper end in px*px: {}'.format(
                            settings['extreme area outliers lower end in px*px'],
                            settings['extreme area outliers upper end in px*px']))
        return None
    if frame_width is None or frame_height is None:
        frame_width = settings['frame width']
        frame_height = settings['frame height']
    if frame_height <= 0 or frame_width <= 0:
        logger.critical('Frame width or frame height 0 or negative; cannot continue. Width: {}, height: {}'.format(
            frame_width, frame_height))
        return None
    if settings['pixel per micrometre'] <= 0:
        logger.critical('\'pixel per micrometre\' setting in tracking.ini 0 or negative. '
                        'Cannot continue. Value: {}'.format(settings['pixel per micrometre']))
        return None
    if not isinstance(df, pd.DataFrame):
        try:
            df = pd.read_csv(path_to_file, sep=',', header=0, usecols=list(DTYPES), dtype=DTYPES)
        except Exception as ex:  # get_data() logs and returns None (helper_file.py:846-919)
            logger.critical('Error reading data frame from file {}: {!r}'.format(path_to_file, ex))
            return None
    params = select_params(settings, fps, frame_height, frame_width)
    good, clean_index, kicks, info = select_rows(df['TRACK_ID'].to_numpy(), df['POSITION_T'].to_numpy(), df['POSITION_X'].to_numpy(),
                                                 df['POSITION_Y'].to_numpy(), df['WIDTH'].to_numpy(), df['HEIGHT'].to_numpy(),
                                                 params, device)
    status = int(info[_lib.SI_STATUS])
    if status == _lib.SEL_TOO_SHORT_BEFORE:
        logger.critical('File is empty/of insufficient length before initial clean-up. '
                        'Minimal size (frames): {}, length: {}, path: {}'.format(params.min_len_frames, df.shape[0], path_to_file))
        return None
    if status == _lib.SEL_TOO_SHORT_AFTER:
        logger.warning('File is empty/of insufficient length after initial clean-up. '
                       'Minimal size: {}, length: {}, path: {}'.format(params.min_len_frames, int(info[_lib.SI_ROWS_AFTER]), path_to_file))
        return None
    n0, n1 = int(info[_lib.SI_TRACKS_BEFORE]), int(info[_lib.SI_TRACKS_AFTER])
    r0, r1 = int(info[_lib.SI_ROWS_BEFORE]), int(info[_lib.SI_ROWS_AFTER])
    logger.info('Tracks before initial cleanup: {}, after: {}, loss: {:.4%}, '
                'data frame entries before: {}, after: {}, loss: {:.4%}'.format(n0, n1, (n0 - n1) / n0, r0, r1, (r0 - r1) / r0))
    if settings['percent quantiles excluded area'] > 0:
        logger.info('Area quartiles: 10%: {:.2f}, 90%: {:.2f}'.format(info[_lib.SI_Q1_AREA], info[_lib.SI_Q3_AREA]))
    if settings['try to omit motility outliers']:
        pct = info[_lib.SI_OUTLIERS] / r1
        logger.info('25/75 % Distance quartiles: {:.3f}, {:.3f} upper outliers: {:.3f} counts: {}, of all entries: {:.4%}'.format(
            info[_lib.SI_Q1_DIST], info[_lib.SI_Q3_DIST], info[_lib.SI_FENCE], int(info[_lib.SI_OUTLIERS]), pct))
        if info[_lib.SI_OUTLIERS_OFF]:
            logger.warning('Motility outliers more than {:.2%} of all data points ({:.2%}); recommend to re-analyse file with '
                           'outlier removal changed if upper quartile is especially low(Quartile: {:.3f})'.format(
                               settings['stop excluding motility outliers if total count above percent'], pct, info[_lib.SI_Q3_DIST]))
            logger.info('Distance outlier exclusion switched off due to too many outliers')
    n_good = int(info[_lib.SI_GOOD_TRACKS])
    logger.info('All tracks before fine selection: {}, left over: {}, difference: {}'.format(n1, n_good, n1 - n_good))
    kick_list = [int(k) for k in kicks]
    kick_string = ('Total: {9}; size < 600: {8}; holes > 6: {7}; distance outlier: {6}; duration 5% over size: {5}; '
                   'area out of bounds: {4}; ratio wrong: {3}; average x/y not within bounds: {2}; '
                   'min/max xy not within screen: {1}; passed: {0}'.format(*kick_list, sum(kick_list)))
    if kick_list[0] < 1000 and kick_list[0] / sum(kick_list) < 0.3:
        logger.warning('Low amount of accepted tracks')
        logger.warning(kick_string)
    else:
        logger.info(kick_string)
    if status == _lib.SEL_NO_TRACKS:
        logger.warning('File {} has no acceptable tracks.'.format(path_to_file))
        return None
    sel = np.flatnonzero(good)
    out = df.iloc[sel][COLUMNS].copy()
    out.insert(0, 'index', clean_index[sel].astype(np.int64))
    out.reset_index(drop=True, inplace=True)
    if settings['store processed .csv file']:
        out.to_csv(os.path.join(results_directory, file_name) + '_selected_data.csv', index=False)
    out.attrs['kick_reasons'] = kick_list
    return out
