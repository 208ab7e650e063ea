"""Import the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

Two places hold it: the source tree /root/reference (authoring container only; ``oracle/make_*golden*.py`` record
the golden vectors from it) and the sourceless byte-compiled tree ``oracle/_ref/t2onet`` that ``oracle/build_ref.py``
makes from it (git-ignored; it travels to the GPU box, where ``tests/test_gpu_actor.py`` runs the reference's own
Actor on the new Executor and ``bench.py --impl reference`` times the reference itself).  ``T2O_REFERENCE_ROOT``
overrides both.  The product never imports this module.

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

_HERE = os.path.dirname(os.path.abspath(__file__))


EXT = '.t2oc'          # oracle/build_ref.py: importlib MAGIC_NUMBER + marshal.dumps(code object)


def _find_root():
    cands = [os.environ.get('T2O_REFERENCE_ROOT'), '/root/reference', os.path.join(_HERE, '_ref', 't2onet')]
    for c in cands:
        if c and (os.path.isfile(os.path.join(c, 'models', 'operators.py')) or
                  os.path.isfile(os.path.join(c, 'models', 'operators' + EXT))):
            return c
    return cands[0] or '/root/reference'


REF_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'operators.py')) or \
        os.path.isfile(os.path.join(REF_ROOT, 'models', 'operators' + EXT))


def is_source_tree():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'operators.py'))


class _CompiledTreeFinder:
    """Import hook for the byte-compiled reference tree (oracle/_ref/t2onet/<package>/<module>.t2oc): the reference's
    top-level packages (models, executors, utils, options, datasets; most of them namespace packages, as in the source
    tree) resolve to the compiled files."""
    TOP = ('models', 'executors', 'utils', 'options', 'datasets')

    def __init__(self, root):
        self.root = root

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if name.split('.')[0] not in self.TOP:
            return None
        rel = os.path.join(self.root, *name.split('.'))
        if os.path.isfile(rel + EXT):
            return importlib.util.spec_from_loader(name, self, origin=rel + EXT)
        if os.path.isfile(os.path.join(rel, '__init__' + EXT)):
            spec = importlib.util.spec_from_loader(name, self, origin=os.path.join(rel, '__init__' + EXT), is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        if os.path.isdir(rel):
            spec = importlib.machinery.ModuleSpec(name, None, is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import importlib.util
        import marshal
        with open(module.__spec__.origin, 'rb') as f:
            data = f.read()
        n = len(importlib.util.MAGIC_NUMBER)
        if data[:n] != importlib.util.MAGIC_NUMBER:
            raise ImportError('%s was compiled by another Python version; rerun python -m oracle.build_ref' % module.__spec__.origin)
        module.__file__ = module.__spec__.origin
        exec(marshal.loads(data[n:]), module.__dict__)


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

    if is_source_tree():
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
    elif not any(isinstance(f, _CompiledTreeFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _CompiledTreeFinder(REF_ROOT))
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


def actor_options(**over):
    """The option namespace experiments/t2onet/train_seq2seqL1.py builds (options/seq2seqGAN_train_options.py defaults),
    with the vocabulary directory pointing into the reference tree."""
    load()
    from options.seq2seqGAN_train_options import TrainOptions
    opt = TrainOptions().parser.parse_args([])
    opt.vocab_dir = os.path.join(REF_ROOT, 'data', 'language')
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def _lengths_compat():
    """models/lang_encoder.py:94 hands pack_padded_sequence the lengths as a tensor on the model's device, which the
    torch the reference was written for accepted; this image's torch (2.11) wants them on the CPU.  Version shim, like
    the import shims above: the lengths are moved, nothing else changes."""
    import torch
    import torch.nn.utils.rnn as rnn
    if getattr(rnn.pack_padded_sequence, '_t2o_compat', False):
        return
    orig = rnn.pack_padded_sequence

    def pack_padded_sequence(input, lengths, *a, **k):
        if isinstance(lengths, torch.Tensor) and lengths.is_cuda:
            lengths = lengths.cpu()
        return orig(input, lengths, *a, **k)
    pack_padded_sequence._t2o_compat = True
    rnn.pack_padded_sequence = pack_padded_sequence


def build_actor(opt, executor_cls=None, seed=10):
    """models/actor.py:Actor of the reference, unmodified.  `executor_cls` stands in for the name `Executor` the module
    imported from executors.executor (models/actor.py:11,49) -- this is the whole switch a user of the reference makes
    (INTEGRATION.md).  The GloVe table (an .h5 file the loader needs h5py for, utils/text_utils.py:70-73) is replaced
    by a seeded random table of the same shape; everything else is the reference's own code and initialisation."""
    import torch
    load()
    _lengths_compat()
    import models.actor as actor_mod
    from utils.text_utils import load_vocab
    n_vocab = len(load_vocab(opt.vocab_dir, opt.dataset, opt.session)[0])

    def load_embedding(path):
        g = torch.Generator().manual_seed(1234)
        return torch.randn(n_vocab - 4, opt.word_vec_dim, generator=g) * 0.3

    actor_mod.load_embedding = load_embedding
    saved = actor_mod.Executor
    if executor_cls is not None:
        actor_mod.Executor = executor_cls
    try:
        torch.manual_seed(seed)
        actor = actor_mod.Actor(opt)
    finally:
        actor_mod.Executor = saved
    return actor
