"""rumpy_b200 -- B200-native (sm_100a) EDSR/RCAN trunk behind RUMpy's model-handler API."""
__version__ = '0.1.0'
