"""B200-native style-transfer hot path (StyleTransferNet + VGG-19 perceptual loss) behind the
reference's Python module / loss-function API.  See DESIGN.md and INTEGRATION.md."""
__all__ = ["engine", "ops"]
