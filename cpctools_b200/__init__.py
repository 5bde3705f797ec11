"""cpctools_b200 -- the SOAP post-processing hot path of GMPavanLab/cpctools (``SOAPify``) on B200.

Drop-in for that path ONLY: expansion of compressed power spectra, normalisation, the SOAP
distance family, timeSOAP and the (all-pairs) distance matrix / classification, flat-exported
under the reference's names like ``src/SOAPify/__init__.py:8-14`` does.  Everything else of
SOAPify (engines, HDF5 storage, LENS, transitions, CLI) is out of scope.  The compute runs in
``libsoapb200.so`` (hand-written sm_100a CUDA, C ABI in ``include/soap_b200.h``); there is no CPU
fallback.
"""
from .utils import *  # noqa: F401,F403
from .distances import *  # noqa: F401,F403
from .analysis import *  # noqa: F401,F403
from .classify import *  # noqa: F401,F403
from .transitions import *  # noqa: F401,F403
from . import utils, distances, analysis, classify, streaming, transitions  # noqa: F401
from ._lib import SoapB200Error, build_library, launch_count  # noqa: F401

__version__ = "0.1.0"
