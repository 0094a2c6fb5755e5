"""Drop-in handler classes for the reference's STT / LLM / TTS slots (S/backend_registry.py:289-503)."""
