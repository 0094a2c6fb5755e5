"""CPU oracle for the speech-to-speech hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain numpy restatement of the arithmetic the
reference delegates to ``transformers`` (Whisper / Llama) and ``scipy`` (TTS
post-processing).  It exists to check the CUDA path; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package
(``speech_to_speech_b200``) never imports from here and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * whisper_ref  -- pinned against transformers 5.5.0 (the package the
    reference's WhisperSTTHandler calls) through ``tests/golden/*.npz`` made by
    ``tests/golden/make_golden.py`` in the builder container.
  * llama_ref    -- pinned the same way against transformers LlamaForCausalLM.
  * tts_post_ref -- pinned against scipy.signal.resample_poly + the reference
    handler's own int16/trim/blocking helpers (pure functions, vectors committed).
  * Qwen3-TTS talker/codec math: PARITY UNPINNED (upstream package absent).
"""
