/* analyzer/analyzer.h -- the suscan analyzer API the reference's Suscan::Analyzer wrapper drives
 * (Suscan/Analyzer.cpp:111-638), as a shim over sdb_analyzer_* (include/sigdigger_b200.h).  Every *_async call
 * returns at once; the reply is a message carrying the caller's req_id.  The analyzer runs on a GPU
 * (suscan_source_config_set_gpu) and fails to start without one: there is no CPU path. */
#ifndef _SUSCAN_ANALYZER_H
#define _SUSCAN_ANALYZER_H
#include <sigutils/types.h>
#include <analyzer/mq.h>
#include <analyzer/msg.h>
#include <analyzer/source.h>
#ifdef __cplusplus
extern "C" {
#endif

struct suscan_analyzer;
typedef struct suscan_analyzer suscan_analyzer_t;

enum suscan_analyzer_sweep_strategy {
  SUSCAN_ANALYZER_SWEEP_STRATEGY_STOCHASTIC = 0, SUSCAN_ANALYZER_SWEEP_STRATEGY_PROGRESSIVE = 1
};
enum suscan_analyzer_spectrum_partitioning {
  SUSCAN_ANALYZER_SPECTRUM_PARTITIONING_DISCRETE = 0, SUSCAN_ANALYZER_SPECTRUM_PARTITIONING_CONTINUOUS = 1
};

typedef SUBOOL (*suscan_analyzer_baseband_filter_func_t)(void *privdata, suscan_analyzer_t *analyzer,
                                                         SUCOMPLEX *samples, SUSCOUNT length, SUSCOUNT offset);

/* library set-up (Suscan/Library.cpp:95-204): registries are static here, the calls only succeed */
#define SUSCAN_MODE_NOLOG 2
SUBOOL suscan_sigutils_init(int mode);
SUBOOL suscan_init_sources(void);
SUBOOL suscan_init_estimators(void);
SUBOOL suscan_init_spectsrcs(void);
SUBOOL suscan_init_inspectors(void);

suscan_analyzer_t *suscan_analyzer_new(const struct suscan_analyzer_params *params, suscan_source_config_t *config,
                                       struct suscan_mq *mq);
void   suscan_analyzer_destroy(suscan_analyzer_t *analyzer);
void  *suscan_analyzer_read(suscan_analyzer_t *analyzer, uint32_t *type);
void  *suscan_analyzer_read_timeout(suscan_analyzer_t *analyzer, uint32_t *type, unsigned int timeout_ms);
void   suscan_analyzer_dispose_message(uint32_t type, void *ptr);
void   suscan_analyzer_req_halt(suscan_analyzer_t *analyzer);
SUSCOUNT suscan_analyzer_get_samp_rate(const suscan_analyzer_t *analyzer);
SUFLOAT  suscan_analyzer_get_measured_samp_rate(const suscan_analyzer_t *analyzer);
void     suscan_analyzer_get_source_time(const suscan_analyzer_t *analyzer, struct timeval *tv);
struct suscan_source_info *suscan_analyzer_get_source_info(const suscan_analyzer_t *analyzer);

SUBOOL suscan_analyzer_set_freq(suscan_analyzer_t *a, SUFREQ freq, SUFREQ lnb);
SUBOOL suscan_analyzer_set_gain(suscan_analyzer_t *a, const char *name, SUFLOAT value);
SUBOOL suscan_analyzer_set_antenna(suscan_analyzer_t *a, const char *name);
SUBOOL suscan_analyzer_set_bw(suscan_analyzer_t *a, SUFLOAT bw);
SUBOOL suscan_analyzer_set_ppm(suscan_analyzer_t *a, SUFLOAT ppm);
SUBOOL suscan_analyzer_set_dc_remove(suscan_analyzer_t *a, SUBOOL remove);
SUBOOL suscan_analyzer_set_iq_reverse(suscan_analyzer_t *a, SUBOOL reverse);
SUBOOL suscan_analyzer_set_agc(suscan_analyzer_t *a, SUBOOL set);
SUBOOL suscan_analyzer_set_hop_range(suscan_analyzer_t *a, SUFREQ min, SUFREQ max);
SUBOOL suscan_analyzer_set_rel_bandwidth(suscan_analyzer_t *a, SUFLOAT rel_bw);
SUBOOL suscan_analyzer_set_buffering_size(suscan_analyzer_t *a, SUSCOUNT size);
SUBOOL suscan_analyzer_set_sweep_stratrgy(suscan_analyzer_t *a, enum suscan_analyzer_sweep_strategy s);   /* sic */
SUBOOL suscan_analyzer_set_spectrum_partitioning(suscan_analyzer_t *a, enum suscan_analyzer_spectrum_partitioning p);
SUBOOL suscan_analyzer_set_history_size(suscan_analyzer_t *a, SUSCOUNT size);
SUBOOL suscan_analyzer_seek(suscan_analyzer_t *a, const struct timeval *pos);
SUBOOL suscan_analyzer_replay(suscan_analyzer_t *a, SUBOOL replay);
SUBOOL suscan_analyzer_set_throttle_async(suscan_analyzer_t *a, SUSCOUNT samp_rate, uint32_t req_id);
SUBOOL suscan_analyzer_set_params_async(suscan_analyzer_t *a, const struct suscan_analyzer_params *params,
                                        uint32_t req_id);
SUBOOL suscan_analyzer_register_baseband_filter(suscan_analyzer_t *a, suscan_analyzer_baseband_filter_func_t fn,
                                                void *privdata);
SUBOOL suscan_analyzer_register_baseband_filter_with_prio(suscan_analyzer_t *a,
                                                          suscan_analyzer_baseband_filter_func_t fn, void *privdata,
                                                          int64_t prio);

SUBOOL suscan_analyzer_open_async(suscan_analyzer_t *a, const char *class_name, const struct sigutils_channel *channel,
                                  uint32_t req_id);
SUBOOL suscan_analyzer_open_ex_async(suscan_analyzer_t *a, const char *class_name,
                                     const struct sigutils_channel *channel, SUBOOL precise, SUHANDLE parent,
                                     uint32_t req_id);
SUBOOL suscan_analyzer_close_async(suscan_analyzer_t *a, SUHANDLE handle, uint32_t req_id);
SUBOOL suscan_analyzer_set_inspector_id_async(suscan_analyzer_t *a, SUHANDLE handle, uint32_t inspector_id,
                                              uint32_t req_id);
SUBOOL suscan_analyzer_set_inspector_config_async(suscan_analyzer_t *a, SUHANDLE handle,
                                                  const suscan_config_t *config, uint32_t req_id);
SUBOOL suscan_analyzer_set_inspector_watermark_async(suscan_analyzer_t *a, SUHANDLE handle, SUSCOUNT watermark,
                                                     uint32_t req_id);
SUBOOL suscan_analyzer_set_inspector_freq_overridable(suscan_analyzer_t *a, SUHANDLE handle, SUFREQ freq);
SUBOOL suscan_analyzer_set_inspector_bandwidth_overridable(suscan_analyzer_t *a, SUHANDLE handle, SUFLOAT bw);
SUBOOL suscan_analyzer_inspector_set_spectrum_async(suscan_analyzer_t *a, SUHANDLE handle, uint32_t spectsrc_id,
                                                    uint32_t req_id);
SUBOOL suscan_analyzer_inspector_estimator_cmd_async(suscan_analyzer_t *a, SUHANDLE handle, uint32_t estimator_id,
                                                     SUBOOL enabled, uint32_t req_id);

#ifdef __cplusplus
}
#endif
#endif
