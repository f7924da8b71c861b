"""Model code of the B200-native SkyEye forward path."""
