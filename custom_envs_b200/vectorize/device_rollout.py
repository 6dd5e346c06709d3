"""Policy in the loop without leaving the GPU.

The reference's runner (run_multiagent_exp_single.py:30-49, stable-baselines PPO2/A2C over
``OptVecEnv``) pulls the [sum(P), 3H] observation matrix to the host every step: 12.5 GB per step at
BASELINE config 4, which caps the host-facing VecEnv at the PCIe rate (~17 k env-steps/s per GPU).
``device_rollout`` keeps the loop on the device: a shared per-agent policy (any callable on a CUDA
tensor; ``SharedMlpPolicy`` has the shape of stable-baselines' ``MlpPolicy``: two tanh layers of 64
for the action mean and for the value) reads the observation rows where the step kernel wrote them
and writes the action vector the next step reads.  Rows are processed in chunks so the hidden
activations of 2e8 agent rows never exist at once.  torch (cuBLAS) runs the policy: it is the
caller's model, not part of the env path."""
import torch


class SharedMlpPolicy(torch.nn.Module):
    """One policy shared by every agent row: obs [rows, 3H] -> (action mean [rows], value [rows])."""

    def __init__(self, obs_dim, hidden=(64, 64), log_std=-1.0):
        super().__init__()
        def tower():
            layers, width = [], obs_dim
            for units in hidden:
                layers += [torch.nn.Linear(width, units), torch.nn.Tanh()]
                width = units
            return torch.nn.Sequential(*layers, torch.nn.Linear(width, 1))
        self.pi, self.vf = tower(), tower()
        self.log_std = torch.nn.Parameter(torch.tensor(float(log_std)))

    def forward(self, obs):
        return self.pi(obs).squeeze(-1), self.vf(obs).squeeze(-1)

    @torch.no_grad()
    def act(self, obs, generator=None):
        mean = self.pi(obs).squeeze(-1).float()
        noise = torch.randn(mean.shape, device=mean.device, generator=generator)
        return mean + noise * self.log_std.exp()


@torch.no_grad()
def policy_actions(policy_act, obs, out, row_chunk=1 << 22, low=-4.0, high=6.0):
    """out[rows] = clip(policy_act(obs[rows]), low, high), chunk by chunk (the action Box of
    MultiOptLRs is [-4, 6], utils_env.get_action_space_optlrs)."""
    rows = obs.shape[0]
    for lo in range(0, rows, row_chunk):
        hi = min(rows, lo + row_chunk)
        out[lo:hi] = policy_act(obs[lo:hi]).clamp_(low, high)
    return out


@torch.no_grad()
def device_rollout(env, policy_act, steps, row_chunk=1 << 22, on_step=None):
    """``steps`` lock-step env steps driven by ``policy_act`` (obs chunk -> action chunk), all on
    ``env.device``.  ``env`` is a ``BatchedOptEnv`` that has been reset.  Returns the per-env
    reward sum [E] (device) and the number of finished episodes; ``on_step(t, obs, reward, done,
    info)`` sees the device tensors of every step (e.g. to fill a PPO rollout buffer)."""
    actions = torch.empty(env.num_rows, dtype=torch.float32, device=env.device)
    returns = torch.zeros(env.num_envs, dtype=torch.float64, device=env.device)
    finished = torch.zeros((), dtype=torch.int64, device=env.device)
    obs = env.obs
    for t in range(steps):
        policy_actions(policy_act, obs, actions, row_chunk)
        obs, reward, done, info = env.step(actions)
        returns += reward
        finished += done.sum()
        if on_step is not None:
            on_step(t, obs, reward, done, info)
    return returns, int(finished.item())
