from planar_optical_flow_b200.dataset import FlowDataset, SyntheticFlowDataset, create_flow_dataloader  # noqa: F401
