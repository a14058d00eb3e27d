"""Process-global ASR model singleton — same functions and error behaviour as
/root/reference/backend/models_manager.py:16-32,52-62 (VAD is out of scope and stays in the reference)."""
import logging
from typing import Optional

from .asr import ASRModel
from .config import AppConfig

logger = logging.getLogger("model_manager")
_asr_model: Optional[ASRModel] = None


def asr_model_init(**kwargs) -> None:
    """Idempotent (models_manager.py:26-32)."""
    global _asr_model
    if _asr_model is not None:
        logger.warning("ASR model already initialized. Skipping re-initialization.")
        return
    kwargs.setdefault("max_batch", AppConfig.SONIC_MAX_BATCH)
    _asr_model = ASRModel(AppConfig.CHECKPOINT_PATH, device=AppConfig.DEVICE, mode=AppConfig.SONIC_MODE, **kwargs)


def asr_model_get() -> ASRModel:
    if _asr_model is None:
        logger.error("Attempted to access uninitialized ASR model")
        raise RuntimeError("ASR model not initialized. Call asr_model_init() first!")
    return _asr_model


def asr_model_reset() -> None:
    """Test helper: drop the singleton (the reference deletes ``asr_model.model`` at shutdown, main.py:84-88)."""
    global _asr_model
    if _asr_model is not None:
        _asr_model.close()
    _asr_model = None
