/*
 * oracle_api.h -- result record shared by the two CPU checkers (TEST INFRASTRUCTURE ONLY):
 *   oracle/_ref/libmcxref.so    the reference's own kernel source compiled for the host, and
 *   oracle/libmcxoracle.so      the plain-C restatement of the same algorithm (mcx_oracle.c).
 * Both take the product's mcxb_config (include/mcxb200.h) so a test can hand one and the same
 * configuration to the CUDA path and to the checkers.
 */
#ifndef MCXB200_ORACLE_API_H
#define MCXB200_ORACLE_API_H

#include <stdint.h>
#include "../include/mcxb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcxo_result {
    /* caller-allocated */
    float*    field;        /* fieldlen floats: primary + shadow halves already folded, NOT normalised */
    uint64_t  fieldlen;     /* in: capacity, out: used */
    float*    energy;       /* 2*nthread floats {escaped, launched} per work-item, or NULL */
    float*    detphoton;    /* detcap*reclen floats or NULL */
    uint64_t* seeddata;     /* detcap*2 or NULL */
    uint32_t  detcap;
    /* outputs */
    uint32_t  detected, reclen;
    double    energytot, energyesc;
    uint64_t  n_segment;    /* hitgrid() calls            */
    uint64_t  n_deposit;    /* fluence atomic adds        */
    uint64_t  n_scatter;    /* scattering-length draws    */
    uint64_t  n_launch;     /* photons launched           */
    double    runtime_ms;
    /* trajectories (MCX_DEBUG_MOVE / MCX_DEBUG_MOVE_ONLY): caller-allocated trajcap * 6 floats or NULL */
    float*    traj;
    uint32_t  trajcap;
    uint32_t  trajcount;    /* out: records stored */
} mcxo_result;

#ifdef __cplusplus
}
#endif
#endif
