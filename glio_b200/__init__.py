"""glio_b200 — B200-native (sm_100a) implementation of GLIO's per-scan LiDAR hot path."""
