"""Run a reference example script UNCHANGED against idrlnet_b200 (aliased as `idrlnet`) for a few
iterations -- build-container check of the "examples run unchanged" boundary:

    python tools/run_reference_example.py /root/reference/examples/ldc/ldc.py [max_iter]

matplotlib / pyevtk are stubbed (not installed here); `max_iter` is clamped and `network_dir`
redirected to gpurun_out/exrun so nothing is written next to the example."""
import os, sys, types, runpy, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ex = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
# stubs for plotting libs
class _Any(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        full = self.__name__ + '.' + k
        if full in sys.modules:
            return sys.modules[full]
        return _anyobj
class _Obj:
    def __call__(self, *a, **kw):
        return _anyobj
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _anyobj
    def __iter__(self):
        return iter([_anyobj, _anyobj])
    def __getitem__(self, i):
        return _anyobj
_anyobj = _Obj()
for m in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.tri', 'matplotlib.cm', 'matplotlib.colors', 'mpl_toolkits', 'mpl_toolkits.axes_grid1', 'mpl_toolkits.mplot3d', 'pyevtk', 'pyevtk.hl']:
    sys.modules[m] = _Any(m)
import idrlnet_b200
import idrlnet_b200.shortcut
sys.modules['idrlnet'] = idrlnet_b200
for sub in ['shortcut', 'solver', 'geo_utils', 'geo_utils.geo', 'geo_utils.geo_obj', 'pde_op.equations', 'pde_op.operator', 'architecture.mlp', 'architecture.layer', 'data', 'net', 'pde', 'pde_op', 'variable', 'node', 'graph', 'callbacks', 'receivers', 'optim', 'architecture', 'header', 'torch_util']:
    try:
        sys.modules['idrlnet.' + sub] = importlib.import_module('idrlnet_b200.' + sub)
    except Exception as e:
        print('no submodule', sub, e)
sc = idrlnet_b200.shortcut
_init = sc.Solver.__init__
def init(self, *a, **kw):
    kw['max_iter'] = min(kw.get('max_iter', 1000), iters)
    kw.setdefault('network_dir', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'exrun', 'net_') + os.path.basename(ex))
    return _init(self, *a, **kw)
sc.Solver.__init__ = init
os.chdir(os.path.dirname(ex))
sys.path.insert(0, os.path.dirname(ex))
try:
    runpy.run_path(ex, run_name='__main__')
    print('EXAMPLE OK', ex)
except SystemExit:
    print('EXAMPLE EXIT', ex)
