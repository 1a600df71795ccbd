"""Actor / critic modules with the reference's parameter names and conventions, so checkpoints interoperate.

Reference: rl/policies/actor.py:142-215 (Gaussian_FF_Actor), rl/policies/critic.py:37-74 (FF_V),
rl/policies/base.py:7-13 (normc_fn).  state_dict keys: actor_layers.{0,1}.{weight,bias}, means.{weight,bias};
critic_layers.{0,1}.{weight,bias}, network_out.{weight,bias}; obs_mean / obs_std are plain attributes.
The modules own the parameters; the rollout and the PPO update read them through `flat_views` (one contiguous
device buffer for actor + critic, so the hand-written kernels and the single gradient all-reduce see one array).
"""
import torch
import torch.nn as nn


def normc_fn(m):
    if m.__class__.__name__.find("Linear") != -1:
        m.weight.data.normal_(0, 1)
        m.weight.data *= 1 / torch.sqrt(m.weight.data.pow(2).sum(1, keepdim=True))
        if m.bias is not None:
            m.bias.data.fill_(0)


class Gaussian_FF_Actor(nn.Module):
    def __init__(self, state_dim, action_dim, layers=(256, 256), env_name=None, fixed_std=None, normc_init=True):
        super().__init__()
        if fixed_std is None:
            raise NotImplementedError("learned std is not on the PPO hot path (ppo.py:536 always passes fixed_std)")
        self.actor_layers = nn.ModuleList([nn.Linear(state_dim, layers[0])] +
                                          [nn.Linear(layers[i], layers[i + 1]) for i in range(len(layers) - 1)])
        self.means = nn.Linear(layers[-1], action_dim)
        self.fixed_std = fixed_std
        self.learn_std = False
        self.action_dim, self.env_name = action_dim, env_name
        self.obs_std, self.obs_mean = 1.0, 0.0
        self.is_recurrent = False
        if normc_init:
            self.apply(normc_fn)
            self.means.weight.data.mul_(0.01)

    def _get_dist_params(self, state):
        x = (state - self.obs_mean) / self.obs_std
        for l in self.actor_layers:
            x = torch.relu(l(x))
        return self.means(x), self.fixed_std

    def forward(self, state, deterministic=True, anneal=1.0):
        mu, sd = self._get_dist_params(state)
        sd = sd * anneal
        return mu if deterministic else torch.distributions.Normal(mu, sd).sample()

    def distribution(self, inputs):
        mu, sd = self._get_dist_params(inputs)
        return torch.distributions.Normal(mu, sd)


class FF_V(nn.Module):
    def __init__(self, state_dim, layers=(256, 256), env_name="NOT SET", normc_init=True, obs_std=None, obs_mean=None):
        super().__init__()
        self.critic_layers = nn.ModuleList([nn.Linear(state_dim, layers[0])] +
                                           [nn.Linear(layers[i], layers[i + 1]) for i in range(len(layers) - 1)])
        self.network_out = nn.Linear(layers[-1], 1)
        self.env_name, self.obs_std, self.obs_mean = env_name, obs_std, obs_mean
        self.is_recurrent = False
        if normc_init:
            self.apply(normc_fn)
        self.train()

    def forward(self, inputs):
        if not self.training:  # critic.py:66 — the critic normalises its input only in eval mode
            inputs = (inputs - self.obs_mean) / self.obs_std
        x = inputs
        for l in self.critic_layers:
            x = torch.relu(l(x))
        return self.network_out(x)


class FF_Actor(nn.Module):
    """rl/policies/actor.py:43-71 — deterministic actor with tanh output (TD3 / DDPG)."""

    def __init__(self, state_dim, action_dim, layers=(256, 256), env_name=None, max_action=1):
        super().__init__()
        self.actor_layers = nn.ModuleList([nn.Linear(state_dim, layers[0])] +
                                          [nn.Linear(layers[i], layers[i + 1]) for i in range(len(layers) - 1)])
        self.network_out = nn.Linear(layers[-1], action_dim)
        self.action_dim, self.env_name, self.max_action = action_dim, env_name, max_action
        self.apply(normc_fn)

    def forward(self, state, deterministic=True):
        x = state
        for l in self.actor_layers:
            x = torch.relu(l(x))
        return torch.tanh(self.network_out(x)) * self.max_action


class Dual_Q_Critic(nn.Module):
    """rl/policies/critic.py:118-168 — two independent Q networks on [state | action]."""

    def __init__(self, state_dim, action_dim, hidden_size=256, hidden_layers=2, env_name="NOT SET"):
        super().__init__()
        self.q1_layers = nn.ModuleList([nn.Linear(state_dim + action_dim, hidden_size)] +
                                       [nn.Linear(hidden_size, hidden_size) for _ in range(hidden_layers - 1)])
        self.q1_out = nn.Linear(hidden_size, 1)
        self.q2_layers = nn.ModuleList([nn.Linear(state_dim + action_dim, hidden_size)] +
                                       [nn.Linear(hidden_size, hidden_size) for _ in range(hidden_layers - 1)])
        self.q2_out = nn.Linear(hidden_size, 1)
        self.env_name = env_name

    def forward(self, state, action):
        x1 = x2 = torch.cat([state, action], -1)
        for l in self.q1_layers:
            x1 = torch.relu(l(x1))
        for l in self.q2_layers:
            x2 = torch.relu(l(x2))
        return self.q1_out(x1), self.q2_out(x2)

    def Q1(self, state, action):
        x1 = torch.cat([state, action], -1)
        for l in self.q1_layers:
            x1 = torch.relu(l(x1))
        return self.q1_out(x1)


def flatten_modules(modules, device):
    """Re-home every parameter of `modules` into one contiguous float32 device buffer (and one for gradients).
    Returns (flat_params, flat_grads, [(name, offset, shape)])."""
    params = [(f"{mi}.{n}", p) for mi, m in enumerate(modules) for n, p in m.named_parameters()]
    total = sum(p.numel() for _, p in params)
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    grad = torch.zeros(total, dtype=torch.float32, device=device)
    index, off = [], 0
    for name, p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view(p.shape)
        p.grad = grad[off:off + n].view(p.shape)
        index.append((name, off, tuple(p.shape)))
        off += n
    return flat, grad, index


def load_reference_checkpoint(path, map_location="cpu"):
    """Load a whole-module checkpoint written by the reference (`torch.save(policy, "actor.pt")`, rl/algos/ppo.py:129-137;
    e.g. trained_models/*/actor.pt) WITHOUT the reference on the import path: the pickle names the classes
    rl.policies.actor.Gaussian_FF_Actor / FF_Actor and rl.policies.critic.FF_V / Dual_Q_Critic, which are resolved to the classes
    of this module (same parameter names, same attributes obs_mean / obs_std / fixed_std)."""
    import pickle
    import sys
    import types

    here = sys.modules[__name__]
    table = {("rl.policies.actor", "Gaussian_FF_Actor"): Gaussian_FF_Actor, ("rl.policies.actor", "FF_Actor"): FF_Actor,
             ("rl.policies.critic", "FF_V"): FF_V, ("rl.policies.critic", "Dual_Q_Critic"): Dual_Q_Critic}

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if (module, name) in table:
                return table[(module, name)]
            if module.startswith("rl.policies"):
                raise pickle.UnpicklingError(f"{module}.{name} is not on the B200 path (feed-forward actors / critics only)")
            return super().find_class(module, name)
    shim = types.ModuleType("apex_b200_ref_pickle")
    shim.Unpickler, shim.load, shim.loads = _Unpickler, pickle.load, pickle.loads
    shim.__name__ = "pickle"
    return torch.load(path, map_location=map_location, pickle_module=shim, weights_only=False)


_REF_CLASS = {"Gaussian_FF_Actor": "rl.policies.actor", "FF_Actor": "rl.policies.actor", "FF_V": "rl.policies.critic",
              "Dual_Q_Critic": "rl.policies.critic"}


def save_reference_checkpoint(module, path):
    """Write `module` the way the reference does — `torch.save(policy, "actor.pt")`, a whole-module pickle (rl/algos/ppo.py:129-137)
    — naming the REFERENCE's classes (rl.policies.actor.Gaussian_FF_Actor, rl.policies.critic.FF_V, ...), so that the reference's
    own tooling (`torch.load` in apex.py:257-280, tools/*) opens it with its own code, and load_reference_checkpoint opens it here.
    The instance carries the attributes the reference's methods read beyond ours (nonlinearity, bounded, action, normc_init, the
    Welford fields of rl/policies/base.py:17-27).  The reference package need not be importable: for the duration of the write
    stand-in modules own those names."""
    import sys
    import types

    name = type(module).__name__
    if name not in _REF_CLASS:
        raise TypeError(f"{name} has no reference counterpart")
    modname = _REF_CLASS[name]
    import copy
    shim_cls = type(name, (nn.Module,), {"__module__": modname})
    inst = shim_cls.__new__(shim_cls)
    host = copy.deepcopy(module).to("cpu")  # the reference saves CPU modules; ours may live in one flat device buffer
    for p_ in host.parameters():
        p_.data = p_.data.clone()            # compact storage: a view would drag the whole flat buffer into the file
        p_.grad = None
    inst.__dict__ = dict(host.__dict__)
    for k, v in list(inst.__dict__.items()):
        if torch.is_tensor(v):
            inst.__dict__[k] = v.detach().cpu().clone()
    extra = {"is_recurrent": False, "welford_state_mean": torch.zeros(1), "welford_state_mean_diff": torch.ones(1), "welford_state_n": 1,
             "nonlinearity": torch.nn.functional.relu, "normc_init": False}
    if name == "Gaussian_FF_Actor":
        extra.update(action=None, bounded=False, learn_std=False)
    if name in ("FF_V", "Dual_Q_Critic"):
        extra.update(welford_reward_mean=0.0, welford_reward_mean_diff=1.0, welford_reward_n=1)
    for k, v in extra.items():
        inst.__dict__.setdefault(k, v)
    saved = {k: sys.modules.get(k) for k in ("rl", "rl.policies", modname)}
    try:
        for k in saved:
            sys.modules[k] = types.ModuleType(k)
        setattr(sys.modules[modname], name, shim_cls)
        torch.save(inst, path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
