/* TEST INFRASTRUCTURE — CPU oracle: TinyRenderer restatement (fp32). See orc_render.cpp. */
#ifndef ORC_RENDER_H
#define ORC_RENDER_H
#include "orc_env.h"
void orc_eye_camera(float* view16, float* proj16);   /* column-major, as pybullet returns them */
void orc_render_eye(const OrcEnv* e, uint8_t* rgb, float* depth, int32_t* mask);
void orc_render_with(const OrcScene* sc, const double* frames /*[S][12]*/, const float* view16, const float* proj16,
                     uint8_t* rgb, float* depth, int32_t* mask);
#endif
