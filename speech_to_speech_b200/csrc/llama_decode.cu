// llama_decode.cu -- persistent Llama-family decode kernel (lands after the Whisper path is parity-green).
#include "common.cuh"
