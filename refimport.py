"""Import the UNMODIFIED reference (skycrapers/TecoGAN-PyTorch) from baseline/_ref/ -- the install
made by tools/vendor_reference.py -- or, in the build container, from /root/reference.

Used only by `bench.py --impl reference`, tests/ and oracle/gen_golden.py; the product package never
imports it.  Recipe = SURVEY.md section 9: no reference file is edited; the modules this image lacks
(skimage / IPython behind metrics/__init__, lmdb behind data/) are stubbed in sys.modules and two
renamed library symbols are aliased.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(ROOT, 'baseline', '_ref', 'codes'), '/root/reference/codes')


def reference_codes_dir():
    for p in CANDIDATES:
        if os.path.isfile(os.path.join(p, 'models', 'networks', 'tecogan_nets.py')):
            return p
    return None


def available():
    return reference_codes_dir() is not None


def _prepare():
    codes = reference_codes_dir()
    if codes is None:
        raise ImportError('reference not installed: run tools/vendor_reference.py in the build container '
                          '(baseline/_ref/ is git-ignored and shipped to the GPU box by gpurun)')
    if codes not in sys.path:
        sys.path.insert(0, codes)
    if 'metrics' not in sys.modules or getattr(sys.modules['metrics'], '__refimport__', None) != codes:
        m = types.ModuleType('metrics')
        m.__path__ = [os.path.join(codes, 'metrics')]    # skip metrics/__init__ (LPIPS -> skimage/IPython)
        m.create_metric_calculator = lambda opt: None    # only needed to import main.py
        m.__refimport__ = codes
        sys.modules['metrics'] = m
    sys.modules.setdefault('lmdb', types.ModuleType('lmdb'))
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, 'gaussian'):            # utils/data_utils.py:15 uses the removed alias
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    return codes


def import_generator():
    """-> (FRNet class, net_utils module, data_utils module) of the reference"""
    _prepare()
    from models.networks.tecogan_nets import FRNet
    from utils import net_utils, data_utils
    return FRNet, net_utils, data_utils


def import_models():
    """-> the reference's `models` package (VSRModel / VSRGANModel / define_generator) and `main`"""
    _prepare()
    import models
    import models.networks
    import models.vsr_model
    import main
    return models, main


def root_dir():
    codes = reference_codes_dir()
    return None if codes is None else os.path.dirname(codes)
