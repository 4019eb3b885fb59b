/*
 * vsr_flat.h — neutral ("flat") interchange form of one VSR.tla state.
 *
 * One VsrFlatState holds the value of all 20 VARIABLES of
 * vsr-revisited/paper/VSR.tla:119-138 with no packing assumptions: message
 * records are spelled out field by field (VSR.tla:157-225, :510-514, :533-541),
 * the message bag is a list of (record, pending-count) pairs (VSR.tla:135), the
 * received-message sets hold whole records (VSR.tla:128-129,133).
 *
 * It is the type the C ABI uses wherever a caller wants to look inside a state
 * (vsr_unpack / vsr_pack / vsr_state_to_tla in vsr_b200.h), and it is the form
 * in which tests hand states between the product and the oracle.  Plain C, no
 * pointers, fixed maximum dimensions.
 */
#ifndef VSR_FLAT_H
#define VSR_FLAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSR_MAX_R 7      /* ReplicaCount upper bound of the flat form */
#define VSR_MAX_V 7      /* |Values| upper bound */
#define VSR_MAX_C 2      /* ClientCount upper bound (the product only accepts 1) */
#define VSR_MAX_MSGS 240 /* |DOMAIN messages| upper bound */

#define VSR_ABSENT 0xFF /* field not part of this record shape */
#define VSR_NIL 0xFE    /* field present with the model value Nil (VSR.tla:849-854) */

/* rep_status values, in VSR.cfg:9-11 order (that order is also TLC's model-value order) */
enum { VSR_NORMAL = 0, VSR_VIEWCHANGE = 1, VSR_RECOVERING = 2 };

/* message type model values, VSR.cfg:12-23 order */
enum {
    VSR_MT_REQUEST = 0,
    VSR_MT_REPLY = 1,
    VSR_MT_PREPARE = 2,
    VSR_MT_PREPAREOK = 3,
    VSR_MT_COMMIT = 4,
    VSR_MT_SVC = 5,
    VSR_MT_DVC = 6,
    VSR_MT_SV = 7,
    VSR_MT_GETSTATE = 8,
    VSR_MT_NEWSTATE = 9,
    VSR_MT_RECOVERY = 10,
    VSR_MT_RECOVERYRESPONSE = 11
};

/* the 19 disjuncts of Next in textual order, VSR.tla:896-918; 0 = Init */
enum {
    VSR_ACT_INIT = 0,
    VSR_ACT_TIMER_SEND_SVC = 1,
    VSR_ACT_RECEIVE_HIGHER_SVC = 2,
    VSR_ACT_RECEIVE_MATCHING_SVC = 3,
    VSR_ACT_SEND_DVC = 4,
    VSR_ACT_RECEIVE_HIGHER_DVC = 5,
    VSR_ACT_RECEIVE_MATCHING_DVC = 6,
    VSR_ACT_SEND_SV = 7,
    VSR_ACT_RECEIVE_SV = 8,
    VSR_ACT_RECEIVE_CLIENT_REQUEST = 9,
    VSR_ACT_RECEIVE_PREPARE = 10,
    VSR_ACT_RECEIVE_PREPARE_OK = 11,
    VSR_ACT_EXECUTE_OP = 12,
    VSR_ACT_SEND_GET_STATE = 13,
    VSR_ACT_RECEIVE_GET_STATE = 14,
    VSR_ACT_RECEIVE_NEW_STATE = 15,
    VSR_ACT_RESTART_EMPTY = 16,
    VSR_ACT_RECEIVES_RECOVERY = 17,
    VSR_ACT_RECEIVES_RECOVERY_RESPONSE = 18,
    VSR_ACT_COMPLETE_RECOVERY = 19,
    VSR_NUM_ACTIONS = 20
};

/* LogEntryType, VSR.tla:157-161.  operation is the 1-based index into Values. */
typedef struct VsrEntry {
    uint8_t view, operation, client, req;
} VsrEntry;

/* Any message record.  Fields a shape does not have are VSR_ABSENT. */
typedef struct VsrMsg {
    uint8_t type;             /* VSR_MT_* */
    uint8_t view, src, dest;  /* view_number, source, dest (replica ids are 1-based as in the spec) */
    uint8_t op, commit;       /* op_number, commit_number */
    uint8_t lnv;              /* last_normal_vn (DVC) */
    uint8_t first_op;         /* first_op (NewState) */
    uint8_t x;                /* x (Recovery, RecoveryResponse) */
    uint8_t has_entry;        /* 1: `message` field present (Prepare) */
    uint8_t has_log;          /* 0: no log field; 1: log is a function on log_lo..log_lo+log_n-1; 2: log = Nil */
    uint8_t log_lo, log_n;    /* sequences have log_lo = 1 */
    uint8_t count;            /* pending deliveries when this record is a key of `messages` */
    uint8_t _pad[2];
    VsrEntry entry;
    VsrEntry log[VSR_MAX_V];
} VsrMsg;

typedef struct VsrClientRow {
    uint8_t req, op, executed, _pad;
} VsrClientRow;

typedef struct VsrReplica {
    uint8_t status, view, op, commit, lnv, sent_dvc, sent_sv, rec_number;
    uint8_t log_n, n_svc, n_dvc, n_rec;
    VsrEntry log[VSR_MAX_V];
    uint8_t peer_op[VSR_MAX_R + 1]; /* [p-1] for peer p */
    VsrClientRow client_table[VSR_MAX_C];
    VsrMsg svc_recv[VSR_MAX_R];
    VsrMsg dvc_recv[VSR_MAX_R];
    VsrMsg rec_recv[VSR_MAX_R];
} VsrReplica;

typedef struct VsrFlatState {
    uint8_t R, C, V;              /* ReplicaCount, ClientCount, |Values| */
    uint8_t aux_svc, aux_restart;
    uint8_t acked[VSR_MAX_V];     /* aux_client_acked[v]: 0 = v not in DOMAIN, 1 = FALSE, 2 = TRUE */
    uint8_t _pad;
    uint16_t n_msgs;
    VsrReplica rep[VSR_MAX_R];
    VsrMsg msgs[VSR_MAX_MSGS];
} VsrFlatState;

#ifdef __cplusplus
}
#endif
#endif /* VSR_FLAT_H */
