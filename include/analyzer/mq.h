/* analyzer/mq.h -- caller-owned message queue (shim).  The reference's wrapper owns a `struct suscan_mq`, initialises
 * it before suscan_analyzer_new and finalises it after suscan_analyzer_destroy (Suscan/MQ.cpp:25-44,
 * Suscan/Analyzer.cpp:601-638): the analyzer posts into it, suscan_analyzer_read drains it.  Blocking MPSC FIFO of
 * (type, payload); the payload's ownership passes to the reader. */
#ifndef _SUSCAN_MQ_H
#define _SUSCAN_MQ_H
#include <sigutils/types.h>
#ifdef __cplusplus
extern "C" {
#endif

struct suscan_mq { void *impl; };

SUBOOL suscan_mq_init(struct suscan_mq *mq);
void   suscan_mq_finalize(struct suscan_mq *mq);
void  *suscan_mq_read(struct suscan_mq *mq, uint32_t *type);                              /* blocks */
void  *suscan_mq_read_w_type(struct suscan_mq *mq, uint32_t type);                        /* blocks until that type */
SUBOOL suscan_mq_poll(struct suscan_mq *mq, uint32_t *type, void **priv);                 /* never blocks */
SUBOOL suscan_mq_timedread(struct suscan_mq *mq, uint32_t *type, void **priv, unsigned int timeout_ms);
SUBOOL suscan_mq_write(struct suscan_mq *mq, uint32_t type, void *priv);
SUBOOL suscan_mq_write_urgent(struct suscan_mq *mq, uint32_t type, void *priv);           /* to the front */

#ifdef __cplusplus
}
#endif
#endif
