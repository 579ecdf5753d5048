from .flow_euler import FlowEulerCfgSampler, FlowEulerGuidanceIntervalSampler, FlowEulerSampler  # noqa: F401
