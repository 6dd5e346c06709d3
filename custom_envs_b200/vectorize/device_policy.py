"""Shared per-agent policy on the device (C ABI: include/b200policy.h; kernel: csrc/b200policy.cu).

The reference's runner evaluates one stable-baselines ``MlpPolicy`` (two tanh layers of 64, linear
head) on every agent row of the observation matrix that ``OptVecEnv.step_wait`` returned
(run_multiagent_exp_single.py:37-49, vectorize/optvecenv.py:78-88).  ``DevicePolicy`` evaluates the
action head of such a network -- taken from a ``SharedMlpPolicy`` / any ``torch.nn.Sequential`` of
that shape -- with bf16 tcgen05 MMAs, either on a dense observation matrix (``act``) or straight on
the env's adjusted-history rings (``act_env``), in which case the env can step ring-only and the
[sum(P), 3H] matrix never exists.  There is no fallback: without the library / an sm_100 GPU the
constructor raises (``SharedMlpPolicy.act`` is the torch path, and it is the caller's choice)."""
import ctypes

import torch

from custom_envs_b200 import _lib

TANH_F32, TANH_BF16X2, TANH_BF16X2_BOTH = 0, 1, 2


def _ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr())


class DevicePolicy:
    def __init__(self, obs_dim, device='cuda:0', tanh_mode=TANH_F32, log_std=None, low=-4.0, high=6.0):
        """``low`` / ``high``: the action Box of MultiOptLRs agents (utils_env.get_action_space_optlrs);
        ``log_std``: log of the diagonal Gaussian's standard deviation, None = deterministic."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.B200EnvError('DevicePolicy needs a CUDA device (no CPU fallback)')
        self.device = torch.device(device)
        self.obs_dim = int(obs_dim)
        self.low, self.high = float(low), float(high)
        self.noise_std = 0.0 if log_std is None else float(torch.as_tensor(log_std).exp())
        self.calls = 0
        self.seed_counter = None
        handle = ctypes.c_void_p()
        if self.lib.b2p_create(self.device.index or 0, self.obs_dim, int(tanh_mode), ctypes.byref(handle)):
            raise _lib.B200EnvError(self.lib.b2p_last_error(None).decode())
        self.handle = handle

    @classmethod
    def from_torch(cls, tower, obs_dim=None, **kwargs):
        """``tower``: Linear(obs_dim, 64), Tanh, Linear(64, 64), Tanh, Linear(64, 1) -- e.g.
        ``SharedMlpPolicy.pi`` -- or a module with such a ``.pi`` (its ``log_std`` is taken along)."""
        if hasattr(tower, 'pi'):
            if 'log_std' not in kwargs and hasattr(tower, 'log_std'):
                kwargs['log_std'] = tower.log_std.detach()
            tower = tower.pi
        linears = [m for m in tower.modules() if isinstance(m, torch.nn.Linear)]
        if len(linears) != 3 or linears[0].out_features != 64 or linears[1].out_features != 64 \
                or linears[2].out_features != 1:
            raise ValueError('DevicePolicy needs an obs_dim -> 64 -> 64 -> 1 tanh tower')
        device = linears[0].weight.device
        policy = cls(obs_dim or linears[0].in_features, device=device, **kwargs)
        policy.set_weights(*[t for lin in linears for t in (lin.weight, lin.bias)])
        return policy

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, code):
        if code:
            raise _lib.B200EnvError(self.lib.b2p_last_error(self.handle).decode())

    def set_weights(self, w1, b1, w2, b2, w3, b3):
        """torch.nn.Linear layouts ([out, in]); the library copies them (fp32), the kernel rounds to bf16."""
        tensors = [t.detach().to(self.device, torch.float32).contiguous() for t in (w1, b1, w2, b2, w3, b3)]
        assert tensors[0].shape == (64, self.obs_dim) and tensors[2].shape == (64, 64) and tensors[4].numel() == 64
        self._check(self.lib.b2p_set_weights(self.handle, *[_ptr(t) for t in tensors], self._stream()))
        torch.cuda.current_stream(self.device).synchronize()      # the sources may be temporaries

    def use_seed_counter(self, enabled=True):
        """Draw the noise seed from a device counter (``self.seed_counter``, int64 [1]) added to the seed argument:
        launches captured in a CUDA graph then see fresh noise on every replay if the graph also advances it."""
        if enabled and self.seed_counter is None:
            self.seed_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        ptr = ctypes.c_void_p(self.seed_counter.data_ptr() if enabled else 0)
        self._check(self.lib.b2p_set_seed_counter(self.handle, ptr))
        if not enabled:
            self.seed_counter = None

    def _seed(self, seed):
        self.calls += 1
        return ctypes.c_uint64((self.calls if seed is None else int(seed)) & (2 ** 64 - 1))

    def act(self, obs, out=None, seed=None):
        """obs [rows, obs_dim] float32 (dense, e.g. ``BatchedOptEnv.obs``) -> actions [rows]."""
        assert obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape[-1] == self.obs_dim
        rows = obs.numel() // self.obs_dim
        if out is None:
            out = torch.empty(rows, dtype=torch.float32, device=self.device)
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == rows
        self._check(self.lib.b2p_act(self.handle, _ptr(obs), rows, _ptr(out), self.noise_std, self._seed(seed),
                                     self.low, self.high, self._stream()))
        return out

    def act_env(self, env, out=None, seed=None):
        """Actions of every agent row of ``env`` (a ``BatchedOptEnv``), read from its rings; VecEnv row order."""
        if out is None:
            out = torch.empty(env.num_rows, dtype=torch.float32, device=self.device)
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == env.num_rows
        self._check(self.lib.b2p_act_env(self.handle, env.handle, _ptr(out), self.noise_std, self._seed(seed),
                                         self.low, self.high, self._stream()))
        return out

    def close(self):
        if getattr(self, 'handle', None):
            torch.cuda.synchronize(self.device)
            self.lib.b2p_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001 (interpreter shutdown)
            pass


def reference_actions(tower, obs, low=-4.0, high=6.0, bf16=True):
    """The same action head in torch, rounding to bf16 where the kernel does (operands of the two
    matmuls; fp32 accumulation, fp32 head): what the parity test holds the kernel to."""
    linears = [m for m in tower.modules() if isinstance(m, torch.nn.Linear)]
    rnd = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if bf16 else (lambda t: t)
    x = rnd(obs.to(torch.float32))
    h1 = torch.tanh(x @ rnd(linears[0].weight.float()).t() + rnd(linears[0].bias.float()))
    h2 = torch.tanh(rnd(h1) @ rnd(linears[1].weight.float()).t() + rnd(linears[1].bias.float()))
    mean = h2 @ linears[2].weight.float().t() + linears[2].bias.float()
    return mean.squeeze(-1).clamp(low, high)


@torch.no_grad()
def device_policy_rollout(env, policy, steps, ring_only=True, on_step=None, use_graph=False):
    """``steps`` lock-step env steps driven by a ``DevicePolicy``; with ``ring_only`` the env never
    writes observation rows and the policy reads the rings (nothing but E rewards / flags / info rows
    and the action vector exists per step).  Returns (per-env reward sums [E], finished episodes).

    ``use_graph``: capture policy -> step -> bookkeeping for two steps (the pipeline's gradient buffers
    alternate) into one CUDA graph and replay it: no per-step Python or launch gaps.  Needs an even
    ``steps``, ``ring_only`` and no ``on_step``; the policy's noise seed then advances on the device."""
    actions = torch.empty(env.num_rows, dtype=torch.float32, device=env.device)
    returns = torch.zeros(env.num_envs, dtype=torch.float64, device=env.device)
    finished = torch.zeros((), dtype=torch.int64, device=env.device)
    obs = env.obs

    def one_step():
        if ring_only:
            policy.act_env(env, actions, seed=0 if use_graph else None)
        else:
            policy.act(obs, actions, seed=0 if use_graph else None)
        out = env.step(actions, ring_only=ring_only)
        returns.add_(out[1])
        finished.add_(out[2].sum())
        return out

    if use_graph:
        if steps % 2 or not ring_only or on_step is not None:
            raise ValueError('use_graph needs an even number of ring-only steps and no on_step hook')
        policy.use_seed_counter(True)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(side):                      # steps 1 and 2 eagerly (lazy module loading outside the capture)
            for _ in range(2):
                one_step()
                policy.seed_counter.add_(1)
        torch.cuda.current_stream(env.device).wait_stream(side)
        torch.cuda.synchronize(env.device)
        with torch.cuda.graph(graph):
            for _ in range(2):
                one_step()
                policy.seed_counter.add_(1)
        for _ in range(steps // 2 - 1):
            graph.replay()
        torch.cuda.synchronize(env.device)
        policy.use_seed_counter(False)
        return returns, int(finished.item())
    for t in range(steps):
        obs, reward, done, info = one_step()
        if on_step is not None:
            on_step(t, obs, reward, done, info)
    return returns, int(finished.item())
