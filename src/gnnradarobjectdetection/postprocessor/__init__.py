"""Reference postprocessor package -> the device-side pieces of radargnn_b200.postprocessor."""
