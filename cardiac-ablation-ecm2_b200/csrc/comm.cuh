// Shared-dof exchange + scalar all-reduce between the ranks of a partitioned mesh (one rank
// per GPU), NCCL over NVLink.  ≙ DeviceConformingProlongationOperator (fem/pfespace.cpp:5259-5532)
// and InnerProduct(comm,...) (linalg/vector.hpp:773-779).  See comm.cu for the design.
#pragma once
#include "common.cuh"

namespace b200pa
{
const unsigned char *comm_owner_mask(b200pa_comm c);
// y[shared dofs] <- sum over all sharing ranks (ascending rank order, identical on every rank)
int comm_exchange_sum(b200pa_comm c, double *yL_dev, const int *done);
int comm_exchange_owner(b200pa_comm c, double *xL_dev);
int comm_allreduce_sum_dev(b200pa_comm c, double *vals_dev, int n);
// step: 1 init, 2 beta, 3 den (PCG scalar steps); *handled = false -> peer path off, nothing was launched
int comm_allreduce_scalar_step(b200pa_comm c, double *val_dev, int step, void *pcg_state, double *norms, bool *handled,
                               const double *extra_dev = nullptr);
bool comm_px(b200pa_comm c);
const unsigned char *comm_shared_mask(b200pa_comm c);
int comm_exchange_sum_apply(b200pa_comm c, double *yL_dev, const int *done, const double *x_dev, const unsigned char *ess_mask,
                            double *dot_out);
// peer path, operator apply in two halves around the segmented reduction: the partial sums of the shared dofs leave for the
// neighbours straight from the slot-order scratch, the receive + finish runs after the reduction (comm.cu)
int comm_px_send_from_slots(b200pa_comm c, const int *offsets_dev, const double *yS_dev, const int *done);
int comm_px_recv_apply(b200pa_comm c, double *yL_dev, const int *done, const double *x_dev, const unsigned char *ess_mask, double *dot_out);
// peer-memory path: device error word raised by a timed-out wait (nullptr when the path is off) and the check the
// host-synchronous entry points run after their final stream synchronisation (nonzero + message when it was raised)
const int *comm_px_err_ptr(b200pa_comm c);
int comm_px_check(b200pa_comm c, const char *where);
// tables set and sized for an L-vector of `ndofs` entries (b200pa_form_set_comm)
int comm_validate(b200pa_comm c, int ndofs);
} // namespace b200pa
