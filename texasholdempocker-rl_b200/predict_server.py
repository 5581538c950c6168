"""Drop-in body for the reference's batch-prediction server process (env/workflow.py:274-593).

The reference runs ``predict_batch_process`` in dedicated processes: game-loop threads put ``(pid, observation)`` on a
shared in-queue and wait for ``(pid, reward_value, action_probs)`` on their out-queue; the server batches what is
waiting and runs either the PyTorch module (fixed ``predict_batch_size``) or one TensorRT engine profile per batch size
(utils/tensorrt_utils.py).  Control traffic rides on five more queues: workflow status changes (acknowledged), model
updates (``update_model_param_queue``: a state_dict in PyTorch mode; acknowledged), and the training hand-shake
(``train_hold_signal_queue`` -> WAITING_MODEL ack -> ``train_update_model_queue`` (step, state_dict) -> two acks).

``predict_batch_process`` below keeps that wire protocol and the fifteen positional parameters, and replaces the compute:
everything that is waiting (up to ``max_batch``) goes through ONE ``BatchedPredictor.predict_batch`` call -- a CUDA-graph
replay per batch bucket on the device that also runs the equity kernels, where the TensorRT path kept one engine profile
per batch size -- and weight updates are applied in place to the live predictor (``load_reference_state_dict``), so no
engine is rebuilt and no graph re-captured.  There is no CPU path: without a CUDA device the predictor raises.

Differences, on purpose: the batch is "what is waiting" in every workflow state (the reference's PyTorch branch waits
for exactly ``predict_batch_size`` rows unless the status is a *_FINISH_WAIT one, its TensorRT branch for the largest
profile that fits); model updates arrive as state_dicts in both modes (there is no .trt file to hand over).
"""
import logging
import os
import queue as _queue

import numpy as np


def _reference_enums():
    """The reference's own status enums (env/constants.py); the server only echoes / compares them."""
    import importlib
    const = importlib.import_module("env.constants")
    return const.WorkflowStatus, const.TrainHoldStatus


def predict_batch_process(in_queue, out_queue_map_dict_train, out_queue_map_dict_eval, batch_predict_model_type, params, update_model_param_queue,
                          workflow_queue, workflow_ack_queue, train_update_model_queue, train_update_model_ack_queue, train_hold_signal_queue,
                          train_hold_signal_ack_queue, predict_batch_out_queue, pid, log_level,
                          interrupt_callback=None, predictor=None, max_batch=1024, dtype=None, enums=None):
    """Same first fifteen parameters as env/workflow.py:274.  ``out_queue_map_dict_*`` and ``batch_predict_model_type`` are
    accepted and unused, as in the reference's live code path (its per-status routing is commented out, :319-326).

    Extras: ``interrupt_callback`` (defaults to the reference's ``tools.interrupt.interrupt_callback`` when importable),
    ``predictor`` (a ready ``BatchedPredictor``; built from ``params['model_param_dict']`` otherwise), ``enums`` =
    (WorkflowStatus, TrainHoldStatus) when the reference's constants are not importable.  Returns the number of rows served."""
    logging.getLogger().setLevel(log_level)
    logging.info(f"predict_batch_process (B200) {pid} pid {os.getpid()} started")
    WorkflowStatus, TrainHoldStatus = enums if enums is not None else _reference_enums()
    if interrupt_callback is None:
        try:
            import importlib
            interrupt_callback = importlib.import_module("tools.interrupt").interrupt_callback
        except Exception:  # noqa: BLE001
            interrupt_callback = lambda: False  # noqa: E731
    if predictor is None:
        import torch
        from .policy_value import BatchedPredictor, PolicyValueNet
        net = PolicyValueNet(**params["model_param_dict"])
        predictor = BatchedPredictor(net, dtype=dtype if dtype is not None else torch.bfloat16, max_batch=max_batch)
        predictor.warmup()
    if hasattr(predict_batch_out_queue, "start_receive_signal_thread"):      # tools/high_performance_queue.py
        predict_batch_out_queue.start_receive_signal_thread()

    workflow_status = WorkflowStatus.DEFAULT
    served = 0
    hold_queues = (train_update_model_queue, train_update_model_ack_queue, train_hold_signal_queue, train_hold_signal_ack_queue)

    def poll_workflow():
        nonlocal workflow_status
        try:
            workflow_status = workflow_queue.get(block=False)
            workflow_ack_queue.put(workflow_status)
        except _queue.Empty:
            pass

    def poll_model_update():
        try:
            state_dict = update_model_param_queue.get(block=False)
        except _queue.Empty:
            return
        predictor.load_reference_state_dict(state_dict)
        logging.info(f"predict_batch_process_{pid} model updated")
        workflow_ack_queue.put(workflow_status)                              # env/workflow.py:422

    def training_hold():
        """env/workflow.py:440-462: the learner asks the servers to wait for its next weights before they predict on."""
        if any(q is None for q in hold_queues):
            return
        try:
            hold_train_step = train_hold_signal_queue.get(block=False)
        except _queue.Empty:
            return
        logging.info(f"predict_batch_process_{pid} received hold signal at training step {hold_train_step}")
        train_hold_signal_ack_queue.put(TrainHoldStatus.WAITING_MODEL)
        while not interrupt_callback():
            try:
                step_num, state_dict = train_update_model_queue.get(block=True, timeout=0.1)
            except _queue.Empty:
                continue
            if state_dict is not None and workflow_status == WorkflowStatus.TRAINING:
                predictor.load_reference_state_dict(state_dict)
                logging.info(f"predict_batch_process_{pid} model training updated at step {step_num}")
            train_update_model_ack_queue.put(WorkflowStatus.TRAINING)
            train_hold_signal_ack_queue.put((TrainHoldStatus.FINISHED_UPDATING, step_num))
            return

    while not interrupt_callback():
        poll_workflow()
        poll_model_update()
        try:
            data_pid, data = in_queue.get(block=True, timeout=0.01)
        except _queue.Empty:
            continue
        pid_list, batch = [data_pid], [data]
        while len(batch) < max_batch:                                         # everything that is waiting: one replay for all of it
            try:
                data_pid, data = in_queue.get(block=False)
            except _queue.Empty:
                break
            pid_list.append(data_pid)
            batch.append(data)
        poll_workflow()
        training_hold()
        action_probs, reward_values = predictor.predict_batch(np.asarray(batch, dtype=np.int32))
        for send_pid, reward_value, probs in zip(pid_list, reward_values, action_probs):
            predict_batch_out_queue.put((send_pid, reward_value, probs))     # env/workflow.py:327
        served += len(batch)
    logging.info(f"predict_batch_process_{pid} detect interrupt after {served} rows")
    return served
