"""Registration of the B200 handlers in the reference's backend registry
(/root/reference/src/speech_to_speech/backend_registry.py:77-100 BackendSpec, :151-161 registry dicts, :206-233
`_simple_handler_factory`).  Two ways in, both leaving the registry module unmodified:

  register()            adds the specs "b200-whisper" (stt) and "b200-transformers" (llm) next to the built-in ones.
                        Must run before `module_arguments` / `s2s_pipeline` freeze the CLI `choices` from the dict
                        keys, i.e. `import speech_to_speech_b200.registry as r; r.register()` first.
  install_overrides()   keeps the names `whisper` / `transformers` and points their lazy factories at our classes
                        (the literal "registry unchanged" reading).
"""
from __future__ import annotations

from typing import Any

STT_MODULE, STT_CLASS = "speech_to_speech_b200.handlers.whisper_stt_handler", "B200WhisperSTTHandler"
LLM_MODULE, LLM_CLASS = "speech_to_speech_b200.handlers.language_model_handler", "B200LanguageModelHandler"


def _registry():
    import importlib
    return importlib.import_module("speech_to_speech.backend_registry")


def register() -> dict[str, Any]:
    br = _registry()
    from speech_to_speech.arguments_classes.whisper_stt_arguments import WhisperSTTHandlerArguments
    specs = {}
    stt = br.BackendSpec(name="b200-whisper", kind="stt", config_type=WhisperSTTHandlerArguments,
                         create_handler=br._simple_handler_factory(STT_MODULE, STT_CLASS, attach_speculative_turns=True),
                         config_prefix="stt")
    br.STT_BACKENDS["b200-whisper"] = stt
    specs["b200-whisper"] = stt
    base = br.LLM_BACKENDS.get("transformers")
    if base is not None:
        import dataclasses

        def _create(ctx: Any, config: Any) -> Any:
            import importlib
            cls = getattr(importlib.import_module(LLM_MODULE), LLM_CLASS)
            return cls(ctx.stop_event, queue_in=ctx.queue_in, queue_out=ctx.queue_out,
                       setup_kwargs={**dict(config), "cancel_scope": ctx.cancel_scope, "speculative_turns": ctx.speculative_turns})

        llm = dataclasses.replace(base, name="b200-transformers", create_handler=_create)
        br.LLM_BACKENDS["b200-transformers"] = llm
        specs["b200-transformers"] = llm
    return specs


def install_overrides() -> None:
    br = _registry()
    import dataclasses
    spec = br.STT_BACKENDS["whisper"]
    br.STT_BACKENDS["whisper"] = dataclasses.replace(
        spec, create_handler=br._simple_handler_factory(STT_MODULE, STT_CLASS, attach_speculative_turns=True))
