from diy_gym_b200.spaces import Box, Dict, Discrete, MultiBinary, MultiDiscrete, Space, Tuple  # noqa: F401
