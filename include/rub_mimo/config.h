/*
 * config.h — drop-in for the reference's mimo/config.h (same macro names and default values,
 * mimo/config.h:35-114).  Every macro can be overridden before inclusion; the receive library
 * itself takes the same parameters at run time through rub_config (rub_mimo.h).
 */
#ifndef RUB_MIMO_CONFIG_H
#define RUB_MIMO_CONFIG_H

/* streamer formats (mimo/config.h:51-52): the offline IQ source uses fc32 files */
#ifndef CPU
#define CPU "fc32"
#endif
#ifndef WIRE
#define WIRE "sc16"
#endif
#ifndef SAMPLING_RATE
#define SAMPLING_RATE 1.0e6
#endif
#ifndef BASEBAND_GAIN
#define BASEBAND_GAIN 0.25
#endif
/* OFDM configuration (mimo/config.h:65-66) */
#ifndef NUM_SUBCARRIERS
#define NUM_SUBCARRIERS 2048
#endif
#ifndef CP_LENGTH
#define CP_LENGTH 152
#endif
/* generator polynomials (mimo/config.h:70-75) */
#define LFSR_SMALL_LENGTH 12
#define LFSR_LARGE_LENGTH 13
#define LFSR_SMALL_0_GEN_POLY 010123
#define LFSR_SMALL_1_GEN_POLY 010151
#define LFSR_LARGE_0_GEN_POLY 020033
#define LFSR_LARGE_1_GEN_POLY 020047
/* misc (mimo/config.h:77-114) */
#ifndef LOG_DIR
#define LOG_DIR "/tmp/"
#endif
#ifndef PLATEAU_THREASHOLD
#define PLATEAU_THREASHOLD 0.95
#endif
#ifndef PID_MAX
#define PID_MAX 1000
#endif
#ifndef USE_ALL_CARRIERS
#define USE_ALL_CARRIERS true
#endif
#ifndef ADD_NULL_CARRIERS
#define ADD_NULL_CARRIERS true
#endif
#ifndef NUM_ACCESS_CODES
#define NUM_ACCESS_CODES 20
#endif
#ifndef NUM_STREAMS
#define NUM_STREAMS 2
#endif
#define BPSK_CONSTELLATION_SIZE 2
#define INVERT_CHANNEL true
#define INVERT_TO_UNITY false
#define SISO false
#define TX_BEAMFORMING 0
/* the reference's LIQUID_MODEM_ARB32OPT / ARITY 32 table is not part of the reference tree;
 * the square Gray QAMs of liquid-dsp are provided instead (bits per symbol) */
#ifndef MODEM_BITS_PER_SYMBOL
#define MODEM_BITS_PER_SYMBOL 2
#endif
#ifndef ARITY
#define ARITY (1 << MODEM_BITS_PER_SYMBOL)
#endif

#endif
