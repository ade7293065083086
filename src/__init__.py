"""Compatibility package: the reference's import paths (`src.utils.utils`,
`src.depracted.model`, ...) resolve to the B200-native modules in planar_optical_flow_b200."""
