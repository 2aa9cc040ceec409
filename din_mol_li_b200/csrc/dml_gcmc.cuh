// dml_gcmc.cuh — grand-canonical insertion/deletion trials on the device (gcmc_run, dana.F90:590-713).
// Included by dml.cu after dml_ctx is defined.
#pragma once

static int gcmc_run_impl(dml_ctx *ctx) {
  FAIL("gcmc_run: not implemented yet");
}
