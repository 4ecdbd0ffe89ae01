"""The tracking.ini keys the hot path reads, with the reference's defaults and the reference's parsing rules.

Mirror of the part of ``ysmr/helper_file.py`` the per-frame loop depends on (create_configs 143-316 for the defaults,
get_configs 586-843 for the conversions).  Rules kept from the reference:

* ``get_configs(x)`` returns ``x`` itself when it is a dict (helper_file.py:595-596) -- the same object, which
  ``track_bacteria`` then MUTATES (track_eval.py:132 flips the sign of 'threshold offset for detection' in place for
  dark-on-light videos).  The drop-in reproduces that mutation.
* an ini path is parsed with configparser, section by section, ``getboolean / getint / getfloat`` like the reference;
* when the reference package is importable its own ``get_configs`` is used, so every key it knows is present for the
  downstream stages (select_tracks, evaluate_tracks).
"""
from __future__ import annotations

import configparser
import logging
import os

# section, key, converter, default  (helper_file.py:160-282)
HOT_PATH_KEYS = [
    ('BASIC RECORDING SETTINGS', 'frames per second', 'float', 30.0),
    ('BASIC RECORDING SETTINGS', 'white bacteria on dark background', 'bool', True),
    ('BASIC RECORDING SETTINGS', 'threshold offset for detection', 'int', 5),
    ('DISPLAY SETTINGS', 'display video analysis', 'bool', True),
    ('RESULTS SETTINGS', 'rename previous result .csv', 'bool', False),
    ('RESULTS SETTINGS', 'delete .csv file after analysis', 'bool', False),
    ('LOGGING SETTINGS', 'verbose', 'bool', False),
    ('ADVANCED VIDEO SETTINGS', 'include luminosity in tracking calculation', 'bool', False),
    ('ADVANCED VIDEO SETTINGS', 'color filter', 'str', 'COLOR_BGR2GRAY'),
    ('ADVANCED VIDEO SETTINGS', 'minimal frame count', 'int', 600),
    ('ADVANCED VIDEO SETTINGS', 'stop evaluation on error', 'bool', True),
    ('ADVANCED VIDEO SETTINGS', 'list save length interval', 'int', 10000),
    ('ADVANCED VIDEO SETTINGS', 'adaptive double threshold', 'float', 2.0),
    ('ADVANCED TRACK DATA ANALYSIS SETTINGS', 'force tracking.ini fps settings', 'bool', False),
    ('GAUSSIAN-SUM FIR FILTER SETTINGS', 'disable gsff', 'bool', False),
    ('GAUSSIAN-SUM FIR FILTER SETTINGS', 'number of LSFFs', 'int', 3),
    ('GAUSSIAN-SUM FIR FILTER SETTINGS', 'minimum horizon size', 'int', 0),
    ('GAUSSIAN-SUM FIR FILTER SETTINGS', 'maximum horizon size', 'int', 30),
    ('TEST SETTINGS', 'debugging', 'bool', False),
]

COLOR_BGR2GRAY = 6  # cv2.COLOR_BGR2GRAY


def default_settings() -> dict:
    d = {key: default for _, key, _, default in HOT_PATH_KEYS}
    d['color filter'] = COLOR_BGR2GRAY
    return d


def _reference_get_configs():
    try:
        from ysmr.helper_file import get_configs  # the reference package, when installed next to us
        return get_configs
    except Exception:
        return None


def get_configs(settings=None):
    """dict -> the same dict (defaults filled in for missing hot-path keys); str/PathLike -> parsed ini; None -> the
    reference looks for ./tracking.ini and we do the same.  Returns None when the file cannot be read, like the
    reference (which also re-creates the file; we never write)."""
    logger = logging.getLogger('ysmr').getChild(__name__)
    if isinstance(settings, dict):
        for key, val in default_settings().items():
            settings.setdefault(key, val)
        return settings
    ref = _reference_get_configs()
    if ref is not None:
        return ref(settings)
    path = os.path.abspath(settings if settings is not None else os.path.join('./', 'tracking.ini'))
    if not os.path.isfile(path):
        logger.critical('tracking.ini not found: %s', path)
        return None
    cp = configparser.ConfigParser(allow_no_value=True)
    cp.read(path)
    out = {}
    try:
        for section, key, kind, default in HOT_PATH_KEYS:
            sec = cp[section]
            if kind == 'bool':
                out[key] = sec.getboolean(key)
            elif kind == 'int':
                raw = sec.get(key)
                if key == 'maximum horizon size':        # helper_file.py:661-667: non-positive / non-int -> None
                    try:
                        out[key] = int(raw)
                        if not out[key] > 0:
                            out[key] = None
                    except (TypeError, ValueError):
                        out[key] = None
                else:
                    out[key] = sec.getint(key)
            elif kind == 'float':
                out[key] = sec.getfloat(key)
            else:
                out[key] = sec.get(key)
        if out['color filter'] == 'COLOR_BGR2GRAY':
            out['color filter'] = COLOR_BGR2GRAY
        assert out['minimum horizon size'] >= 0 and out['number of LSFFs'] > 1 and out['frames per second'] > 0
    except (TypeError, ValueError, KeyError, AssertionError) as ex:
        logger.exception('could not read %s: %r', path, ex)
        return None
    for key, val in out.items():
        if val is None and key != 'maximum horizon size':
            logger.critical('tracking.ini is missing a value in %s', key)
            return None
    out['tracking_ini_filepath'] = path
    return out
