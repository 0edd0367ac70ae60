"""Host-side metadata of the MI metrics: names, declared bounds and scales.

Mirrors the table SENDER_MI_METRICS and get_min_obs_vector / get_max_obs_vector of the reference
(common/sender_obs.py:95-108, 193-206).  The metric VALUES are computed on the device
(csrc/pcc_core.cuh: mi_stats); this module only provides what observation_space needs.
"""
import numpy as np

# (name, min_val, max_val, scale) in the reference's order; the index is the device metric id
METRICS = [
    ("send rate", 0.0, 1e9, 1e7),
    ("recv rate", 0.0, 1e9, 1e7),
    ("recv dur", 0.0, 100.0, 1.0),
    ("send dur", 0.0, 100.0, 1.0),
    ("avg latency", 0.0, 100.0, 1.0),
    ("loss ratio", 0.0, 1.0, 1.0),
    ("ack latency inflation", -1.0, 10.0, 1.0),
    ("sent latency inflation", -1.0, 10.0, 1.0),
    ("conn min latency", 0.0, 100.0, 1.0),
    ("latency increase", 0.0, 100.0, 1.0),
    ("latency ratio", 1.0, 10000.0, 1.0),
    ("send ratio", 0.0, 1000.0, 1.0),
]
METRIC_NAMES = [m[0] for m in METRICS]
DEFAULT_FEATURES = "sent latency inflation,latency ratio,send ratio"  # network_sim.py:348-351


def feature_names(features):
    if isinstance(features, str):
        return features.split(",")
    return list(features)


def feature_ids(features):
    ids = []
    for name in feature_names(features):
        if name not in METRIC_NAMES:
            raise KeyError(name)  # the reference raises KeyError from its metric dict too
        ids.append(METRIC_NAMES.index(name))
    return ids


def get_min_obs_vector(features):
    return np.array([METRICS[i][1] for i in feature_ids(features)])


def get_max_obs_vector(features):
    return np.array([METRICS[i][2] for i in feature_ids(features)])
