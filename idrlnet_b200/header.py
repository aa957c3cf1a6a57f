"""Package-wide constants and the logger (mirror of idrlnet/header.py)."""
import logging
import os

DIFF_SYMBOL = "__"  # idrlnet/header.py:6 -- "u__x__x" is d2u/dx2

logger = logging.getLogger("idrlnet_b200")
if not logger.handlers:
    _fmt = logging.Formatter("[%(asctime)s] [%(levelname)s] %(message)s", datefmt="%d-%b-%y %H:%M:%S")
    _h = logging.StreamHandler()
    _h.setFormatter(_fmt)
    logger.addHandler(_h)
    if os.environ.get("IDRLNET_B200_TRAIN_LOG", "0") == "1":  # the reference always writes ./train.log
        _f = logging.FileHandler("train.log", mode="a")
        _f.setFormatter(_fmt)
        logger.addHandler(_f)
    logger.setLevel(os.environ.get("IDRLNET_B200_LOGLEVEL", "INFO"))
    logger.propagate = False
