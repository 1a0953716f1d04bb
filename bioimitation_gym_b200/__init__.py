"""B200-native batched physics backend for the bioimitation-gym envs."""
