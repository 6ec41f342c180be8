"""TEST STUB of stable_baselines3.common.vec_env.VecEnv (see the package docstring)."""
from abc import ABC, abstractmethod


class VecEnv(ABC):
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]
        render_modes = self.get_attr("render_mode")
        assert all(m == render_modes[0] for m in render_modes)
        self.render_mode = render_modes[0]
        self.metadata = {"render_modes": [] if self.render_mode is None else [self.render_mode]}

    @abstractmethod
    def reset(self): ...

    @abstractmethod
    def step_async(self, actions): ...

    @abstractmethod
    def step_wait(self): ...

    @abstractmethod
    def close(self): ...

    @abstractmethod
    def get_attr(self, attr_name, indices=None): ...

    @abstractmethod
    def set_attr(self, attr_name, value, indices=None): ...

    @abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs): ...

    @abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None): ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def seed(self, seed=None):
        if seed is None:
            seed = 0
        self._seeds = [seed + idx for idx in range(self.num_envs)]
        return self._seeds

    @property
    def unwrapped(self):
        return self


class VecEnvWrapper(VecEnv):
    """What DQN._wrap_env / VecFrameStack build on: takes a VecEnv and forwards."""

    def __init__(self, venv, observation_space=None, action_space=None):
        assert isinstance(venv, VecEnv), "SB3 would wrap a non-VecEnv in DummyVecEnv([lambda: env])"
        self.venv = venv
        super().__init__(venv.num_envs, observation_space or venv.observation_space, action_space or venv.action_space)

    def reset(self):
        return self.venv.reset()

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        return self.venv.step_wait()

    def close(self):
        return self.venv.close()

    def get_attr(self, attr_name, indices=None):
        return self.venv.get_attr(attr_name, indices)

    def set_attr(self, attr_name, value, indices=None):
        return self.venv.set_attr(attr_name, value, indices)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return self.venv.env_method(method_name, *method_args, indices=indices, **method_kwargs)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return self.venv.env_is_wrapped(wrapper_class, indices=indices)
