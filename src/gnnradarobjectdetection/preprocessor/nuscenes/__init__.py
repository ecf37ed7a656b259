"""Hot-path part of reference preprocessor/nuscenes."""
