"""Test alias of the navigation-graph stand-in environment (speaker_follower_b200/navgraph_env.py)."""
from speaker_follower_b200.navgraph_env import FakeImageFeatures, FakeR2RBatch, WorldState  # noqa: F401
