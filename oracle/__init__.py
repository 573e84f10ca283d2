"""CPU oracle for the hot path behind Engine::transcribe_audio (reference src/engine.rs:65-200).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (whisper-diarize-rs_b200/, host/) never
imports, links or executes anything in this package.

PARITY UNPINNED: the reference's arithmetic lives in un-vendored dependencies (whisper.cpp via
whisper-rs 0.15.0, ONNX Runtime + kaldi-native-fbank via pyannote-rs 0.3.1) and its own tests pin
nothing on this path.  The restatements follow SURVEY Appendix A and are
cross-checked against `transformers` / `torchaudio` where the semantics coincide.
"""
