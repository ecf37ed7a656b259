"""Hot-path part of reference preprocessor/radarscenes."""
