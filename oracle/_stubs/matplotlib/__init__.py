"""Stub `matplotlib` (absent from this image): the reference env imports it at
module top (`uav_env.py:10-12`) for rendering that the hot path never calls."""
