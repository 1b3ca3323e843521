"""B200-native photometric Gauss-Newton frame-to-keyframe tracker (the ELLC hot path)."""
