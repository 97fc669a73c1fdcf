"""Stand-in tensorflow.keras.regularizers (training-time only; nothing to do at inference)."""


def l2(l2=0.01, **kw):
    return None
