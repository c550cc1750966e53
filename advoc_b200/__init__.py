"""advoc_b200: B200-native (sm_100a) hot path of the adversarial vocoder.

Modules mirror the reference package: `spectral` (advoc/spectral.py), `audioio`
(advoc/audioio.py), `model` (models/advoc/*.py), with `nets` as the execution engine and
`_native` as the ctypes binding of libadvoc_b200.so.
"""
__version__ = '0.1.0'
