"""TEST STUB: the abstract surface of stable-baselines3 2.7's VecEnv (the reference pins SB3 2.7.0, uv.lock:1651),
written from its documented interface so that tests can check `UAVVecEnv` satisfies it.  SB3 itself is absent from
the build image; this is not SB3 and implements nothing beyond what VecEnv.__init__ / step / seed do."""
