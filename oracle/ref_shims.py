"""Import the UNMODIFIED reference from /root/reference -- TEST INFRASTRUCTURE ONLY.

Only usable in the authoring container (the GPU box has no /root/reference); used
by ``oracle/make_golden.py`` to record golden vectors and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).
Nothing in the product, ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this.

The reference needs three import shims and one class stub:

1. ``kornia``  -> ``oracle.hsv`` (kornia is not installed and not installable here)
2. ``pyutils.edgeconnect.src.{config,edge_connect}`` -> empty classes (the inpainting
   GAN; needs skimage/matplotlib/weights; out of scope)
3. ``h5py`` -> empty module (only imported by utils/text_utils.py:7)
4. ``models.operators.InpaintOperator`` -> identity stub, installed before
   ``executors.executor`` is imported (its real ctor copies files into the
   read-only tree and loads weights, models/operators.py:631-649)
"""
import os
import sys
import types

REF_ROOT = os.environ.get('T2O_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'operators.py'))


_loaded = None


def load():
    """Returns a namespace with the reference modules: .operators .executor .beam_search
    .beam_search_fixed_order .beam_search_eps_greedy .options"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    from . import hsv

    kornia = types.ModuleType('kornia')
    kornia.rgb_to_hsv = hsv.rgb_to_hsv
    kornia.hsv_to_rgb = hsv.hsv_to_rgb
    sys.modules.setdefault('kornia', kornia)

    for name in ['pyutils', 'pyutils.edgeconnect', 'pyutils.edgeconnect.src',
                 'pyutils.edgeconnect.src.config', 'pyutils.edgeconnect.src.edge_connect']:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['pyutils.edgeconnect.src.config'].Config = type('Config', (), {})
    sys.modules['pyutils.edgeconnect.src.edge_connect'].EdgeConnect = type('EdgeConnect', (), {})
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            sys.modules['h5py'] = types.ModuleType('h5py')

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import models.operators as operators

    class _InpaintStub(operators.Operator):
        def __init__(self, cfg):
            super().__init__(cfg)
            self.short_name = 'inpaint_obj'
            self.num_op_param = 1
            self.setup()

        def get_param_range(self):
            return 0, 0, 0

        def process(self, img, param):
            return img

    operators.InpaintOperator = _InpaintStub
    import executors.executor as executor
    import utils.beam_search as beam_search
    import utils.beam_search_fixed_order as beam_search_fixed_order
    import utils.beam_search_eps_greedy as beam_search_eps_greedy
    from options.fiveK_base_options import BaseOptions
    for m in (beam_search, beam_search_fixed_order, beam_search_eps_greedy):
        m.device = 'cpu'
    _loaded = types.SimpleNamespace(
        operators=operators, executor=executor, beam_search=beam_search,
        beam_search_fixed_order=beam_search_fixed_order,
        beam_search_eps_greedy=beam_search_eps_greedy,
        options=lambda: BaseOptions().parser.parse_args([]))
    return _loaded
