// empty stand-in: tests/emul/cune_emul.cpp supplies the few CUDA names the emulated header uses
