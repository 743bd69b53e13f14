"""Model lookup with the reference's search order (`diy_gym/model.py:8,59-63`).

Order: the path as given (cwd / absolute), this package's `data/` (live URDFs), the stand-ins for the
absent `pybullet_data` package (`data/standin/`), any directory in `$DIYGYM_URDF_PATH`, and finally the
pre-compiled descriptors of the reference's vendored URDFs (`data/compiled/<rel>.json`, produced by
`tools/compile_assets.py`).  A miss raises the reference's `ValueError('Could not find URDF: ...')`.
"""
import os

from .compiler.urdf import compile_urdf, load_model

_HERE = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(_HERE, 'data')
STANDIN_DIR = os.path.join(DATA_DIR, 'standin')
COMPILED_DIR = os.path.join(DATA_DIR, 'compiled')

_cache = {}


def urdf_search_path():
    extra = [p for p in os.environ.get('DIYGYM_URDF_PATH', '').split(os.pathsep) if p]
    return ['', DATA_DIR, STANDIN_DIR] + extra


def resolve_model(urdf):
    """Return the model descriptor for a config's `model:` value."""
    if urdf in _cache:
        return _cache[urdf]
    desc = None
    for base in urdf_search_path():
        path = os.path.join(base, urdf)
        if os.path.isfile(path):
            desc = compile_urdf(path, rel_name=urdf)
            break
    if desc is None:
        compiled = os.path.join(COMPILED_DIR, urdf + '.json')
        if os.path.isfile(compiled):
            desc = load_model(compiled)
    if desc is None:
        raise ValueError('Could not find URDF: ' + urdf)
    _cache[urdf] = desc
    return desc
