/* stand-in for <bgfx/embedded_shader.h>: vg.cpp only builds a table of its eight shaders (vg.cpp:24-61) */
#ifndef BGFX_STUB_EMBEDDED_SHADER_H
#define BGFX_STUB_EMBEDDED_SHADER_H
#include "bgfx.h"
#define BGFX_EMBEDDED_SHADER(_name) { #_name }
#define BGFX_EMBEDDED_SHADER_END() { nullptr }
#endif
