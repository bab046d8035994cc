"""Import the reference's own Python modules from /root/reference (this container only).

TEST INFRASTRUCTURE.  Used by the golden-vector generators under tests/golden/ and by
CPU tests that are skipped when /root/reference is absent (it does not exist on the
GPU box).  The reference imports matplotlib and openai-clip unconditionally
(all_utils/utils.py:5,18); neither is installed, so empty stand-ins are registered
first -- they are never called on the paths we exercise.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "all_utils"))


def _stub(name: str):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    return sys.modules[name]


def import_reference_utils():
    """Returns the reference's ``all_utils.utils`` module."""
    if not available():
        raise ImportError("/root/reference is not present")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    clip = _stub("clip")
    clipclip = _stub("clip.clip")
    clip.clip = clipclip
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import all_utils.utils as ref_utils  # noqa

    return ref_utils


def import_reference_cal():
    """Returns the reference's ``fgvc.models.cal`` module (WSDAN_CAL)."""
    if not available():
        raise ImportError("/root/reference is not present")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fgvc.models.cal as cal  # noqa

    return cal
