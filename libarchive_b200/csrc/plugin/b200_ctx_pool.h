/* b200_ctx_pool.h — see b200_ctx_pool.c */
#ifndef B200_CTX_POOL_H
#define B200_CTX_POOL_H
#include "b200inflate.h"
int  b200_ctx_acquire(b2i_ctx **out);   /* B2I_OK or the b2i_ctx_create error */
int  b200_ctx_acquire_dev(int device, b2i_ctx **out);
void b200_ctx_release(b2i_ctx *c, int healthy);
/* pinned host buffer of at least `need` bytes (*cap = its real size), kept for reuse on release */
void *b200_buf_acquire(size_t need, size_t *cap);
void  b200_buf_release(void *p, size_t cap);
/* the same for callers that do not keep the capacity */
void *b200_buf_acquire_tagged(size_t need);
void  b200_buf_release_tagged(void *p);
#endif
