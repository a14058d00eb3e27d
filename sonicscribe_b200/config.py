"""Application constants the hot path needs — same names and defaults as /root/reference/backend/config.py:9-43
(read from the environment the same way; python-dotenv is optional here)."""
import os

try:  # the reference calls load_dotenv() unconditionally (config.py:7)
    from dotenv import load_dotenv

    load_dotenv()
except Exception:  # pragma: no cover
    pass


class AppConfig:
    CHECKPOINT_PATH = os.getenv("CHECKPOINT_PATH", "./checkpoint")
    DEVICE = os.getenv("DEVICE", "cuda")
    AUDIO_SAMPLE_RATE = 16000
    AUDIO_CHUNK_DURATION_MS = 64
    AUDIO_CHUNK_SIZE = int(AUDIO_SAMPLE_RATE * 2 * AUDIO_CHUNK_DURATION_MS / 1000)
    TEMPORARY_TRANSCRIPTION_INTERVAL = 20
    MAX_SEGMENT_DURATION = 30.0
    MAX_SPEECH_SEGMENTS = 3
    # additions of this implementation: never rename the reference's keys above
    SONIC_MODE = os.getenv("SONIC_MODE", "native")          # native | int8 | fp32
    SONIC_MAX_BATCH = int(os.getenv("SONIC_MAX_BATCH", "16"))   # segments the dynamic batcher may coalesce into one device pass
    SONIC_BATCH_WINDOW_MS = float(os.getenv("SONIC_BATCH_WINDOW_MS", "3"))
