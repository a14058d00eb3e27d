"""bytes(int16 LE, 16 kHz) -> text.  Same class, methods, thresholds and error swallowing as
/root/reference/backend/transcription_manager.py:16-65."""
import logging
import traceback

import numpy as np
import torch

from .config import AppConfig
from .models_manager import asr_model_get

logger = logging.getLogger("speech-to-text")


class TranscriptionManager:
    async def transcribe_temporary(self, audio_data: bytes) -> str:
        if not audio_data or len(audio_data) < AppConfig.AUDIO_CHUNK_SIZE:
            return ""
        try:
            return await self._transcribe(audio_data, is_final=False, max_new_tokens=15)
        except Exception as e:  # the reference degrades to "" (transcription_manager.py:26-28)
            logger.error(f"temporary transcription failed: {e}\n{traceback.format_exc()}")
            return ""

    async def transcribe_committed(self, audio_data: bytes, segment_duration: float) -> str:
        if not audio_data or len(audio_data) < AppConfig.AUDIO_CHUNK_SIZE * 2:
            return ""
        try:
            max_new_tokens = min(50 + int(segment_duration * 5), 200)
            return await self._transcribe(audio_data, is_final=True, max_new_tokens=max_new_tokens)
        except Exception as e:
            logger.error(f"committed transcription failed: {e}\n{traceback.format_exc()}")
            return ""

    async def _transcribe(self, audio_data: bytes, is_final: bool, max_new_tokens: int) -> str:
        audio_array = np.frombuffer(audio_data, dtype=np.int16)
        if len(audio_array) == 0:
            return ""
        asr_model = asr_model_get()
        if hasattr(asr_model, "transcribe_pcm16"):
            # same numbers as the reference's int16 -> float32 / 32768 -> [1, N] detour (transcription_manager.py:45-54), with the
            # widening done on the device: the int16 bytes go to the GPU as they arrived from the WebSocket
            result = asr_model.transcribe_pcm16(audio_array, max_new_tokens=max_new_tokens)
            return result.strip()
        audio_tensor = torch.from_numpy(audio_array.copy()).float() / 32768.0
        if audio_tensor.dim() == 1:
            audio_tensor = audio_tensor.unsqueeze(0)
        result = asr_model.transcribe(audio_tensor, sampling_rate=16000, max_new_tokens=max_new_tokens)
        return result.strip()
