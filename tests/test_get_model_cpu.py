"""CPU: the model factory ``diffdock_b200.utils.get_model`` against the class choice and constructor keywords the UNMODIFIED
reference ``get_model`` (utils/utils.py:172-281) produced for six ``model_parameters.yml``-style namespaces
(tests/golden/ref_get_model.pt, tests/golden/make_golden_get_model.py)."""
from argparse import Namespace

import pytest
import torch

from tests.parity_helpers import load_golden


@pytest.mark.parametrize('i', range(6))
def test_get_model_passes_the_reference_keywords(monkeypatch, i):
    from diffdock_b200 import utils as U
    c = load_golden('ref_get_model.pt')[i]
    made = {}

    class Recorder:
        def __init__(self, **kw):
            made.update(kw)

        def to(self, device):
            made['_moved_to'] = device
            return self

    monkeypatch.setattr(U, '_model_class', lambda name: made.setdefault('_class', name) and Recorder)
    m = U.get_model(Namespace(**c['args']), 'cpu', t_to_sigma='T2S', no_parallel=True, **c['call'])
    assert isinstance(m, Recorder) and made.pop('_class') == c['class'] and made.pop('_moved_to') == torch.device('cpu')
    assert made.pop('t_to_sigma') == 'T2S' and made.pop('device') == torch.device('cpu')
    emb = made.pop('timestep_emb_func')
    assert torch.allclose(emb(torch.tensor([0.0, 0.3, 1.0])), c['emb_of_t'], atol=0, rtol=0) or c['args'].get('embedding_type') == 'fourier'
    assert made == c['kwargs'], {k: (made.get(k), c['kwargs'].get(k)) for k in set(made) | set(c['kwargs'])
                                 if made.get(k) != c['kwargs'].get(k)}


def test_constructor_signatures_accept_every_keyword():
    """Every keyword the factory can produce is a parameter of the class it goes to (the classes need CUDA to be built, so the
    signatures are inspected instead)."""
    import inspect
    from diffdock_b200 import utils as U
    for c in load_golden('ref_get_model.pt'):
        name, kw = U.model_kwargs(Namespace(**c['args']), **c['call'])
        params = inspect.signature(U._model_class(name).__init__).parameters
        missing = [k for k in kw if k not in params]
        assert not missing, (name, missing)


def test_get_model_refuses_data_parallel_and_missing_flags():
    from diffdock_b200 import utils as U
    c = load_golden('ref_get_model.pt')[0]
    with pytest.raises(NotImplementedError, match='DataParallel'):
        U.get_model(Namespace(**c['args']), 'cuda', t_to_sigma=None, no_parallel=False)
    a = dict(c['args'])
    del a['ns']
    with pytest.raises(AttributeError, match="'ns'"):
        U.model_kwargs(Namespace(**a))
