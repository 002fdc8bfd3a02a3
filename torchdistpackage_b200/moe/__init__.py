from .layer import MoELayer, TopKGate, Experts
