"""Import the UNMODIFIED reference (read-only at /root/reference) in the build container.

TEST INFRASTRUCTURE.  Used by tests/golden/make_golden.py (to generate the committed golden
vectors) and by tests that compare the oracle with the live reference.  /root/reference does
not exist on the GPU box; ``available()`` is False there and those tests skip.
"""
import os
import sys
import warnings

REF_ROOT = os.environ.get('B200W_REFERENCE_ROOT', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'pytorch_wavelets'))


def load():
    """Return the reference's ``pytorch_wavelets`` module (imports it on first call)."""
    if not available():
        raise RuntimeError('reference not present at %s' % REF_ROOT)
    for p in (_REPO, os.path.join(_HERE, 'pywt_standin'), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import pytorch_wavelets  # noqa: the reference package
    assert os.path.realpath(pytorch_wavelets.__file__).startswith(os.path.realpath(REF_ROOT))
    return pytorch_wavelets
