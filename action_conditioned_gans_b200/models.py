"""`models.py` surface of the reference (models.py:8-88): build_generator, build_generator_transform and
build_discriminator as functions of CUDA tensors.

TF's variable scopes are mirrored by a module-level variable table keyed by scope ('g' / 'd'): the first call
creates the variables (xavier-uniform weights, zero beta / biases: slim defaults) unless reuse=True, a second call
without reuse raises ValueError like tf.get_variable does.  `actions` may be the tiled map the reference feeds
([B,4,4,10]; [B,16,16,10] for the discriminator after repair R3) or the raw [B,10] vector.

The functions are differentiable: they call the `acg::generator_transform` / `acg::generator` / `acg::discriminator`
custom ops (torch_ops.py), whose backward is the engine's hand-scheduled backward pass.  `trainable(scope)` returns
the scope's flat parameter tensor (a leaf sharing storage with the ParamStore) -- the analogue of
`tf.get_collection(TRAINABLE_VARIABLES, scope)` at train.py:87-88 -- and `variables(scope)` its named views.
"""
import numpy as np
import torch

from . import engine as E
from . import torch_ops as T

VARIABLES = {}     # scope -> ParamStore
_LEAVES = {}       # scope -> flat parameter leaf (requires_grad, same storage as the store)
_SEED = 7          # train.py:14


def reset_default_graph():
    VARIABLES.clear()
    _LEAVES.clear()


def trainable(scope):
    """Flat fp32 parameter tensor of scope 'g' / 'd' (requires_grad; `.grad` is filled by backward())."""
    return _LEAVES[scope]


def variables(scope):
    """{tf variable name: view into the flat buffer} of a scope (HWIO weights, beta / biases)."""
    return VARIABLES[scope].views


def _actions(a):
    a = a.float()
    if a.dim() == 4:
        a = a[:, 0, 0, :]
    return a.contiguous()


def _store(scope, spec, reuse, device):
    if scope in VARIABLES:
        if not reuse:
            raise ValueError("Variable %s/conv1/weights already exists, disallowed. Did you mean to set reuse=True?"
                             % scope)
        st = VARIABLES[scope]
        if [L.name for L in st.spec] != [L.name for L in spec] or st.spec != spec:
            raise ValueError("scope '%s' was created with a different architecture" % scope)
        return st
    if reuse:
        raise ValueError("Variable %s/conv1/weights does not exist (reuse=True)" % scope)
    rng = np.random.RandomState(_SEED + len(VARIABLES))
    st = E.ParamStore(spec, device, E.xavier_init(spec, rng))
    VARIABLES[scope] = st
    T.register_store(st)
    _LEAVES[scope] = st.flat.detach().requires_grad_(True)      # same storage: optimizer kernels update it in place
    return st


def _need_cuda(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("acg_b200 models need CUDA tensors: there is no CPU fallback")
    return t.contiguous().float()


def build_generator(images, actions, reuse=False):
    """models.py:8-22 -> tanh image [B,64,64,3]."""
    images = _need_cuda(images)
    _store("g", E.g_direct_spec(), reuse, images.device)
    return T.generator(images, _actions(actions).to(images.device), _LEAVES["g"])


def build_generator_transform(images, actions, batch_size, reuse=False, color_channels=3, ksize=5):
    """models.py:24-74 -> (frame [B,64,64,3], state [B,5])."""
    images = _need_cuda(images)
    if color_channels != 3 or images.shape[0] != batch_size:
        raise ValueError("build_generator_transform: color_channels must be 3 and batch_size must match images")
    _store("g", E.g_dna_spec(ksize), reuse, images.device)
    out, state = T.generator_transform(images, _actions(actions).to(images.device), _LEAVES["g"], ksize)
    return out, state


def build_discriminator(inputs, actions, reuse=False):
    """models.py:76-88: inputs = concat([frame_t, frame_t+1], 3) [B,64,64,6] -> logits [B,2,2,1]."""
    inputs = _need_cuda(inputs)
    _store("d", E.d_spec(), reuse, inputs.device)
    return T.discriminator(inputs, _actions(actions).to(inputs.device), _LEAVES["d"])
