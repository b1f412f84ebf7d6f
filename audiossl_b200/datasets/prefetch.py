"""Host -> device staging of raw-waveform batches (device half): the reference ships
CPU-computed mels through DataLoader workers; here the mel runs on the GPU, so what crosses PCIe is the waveform
(640 KB per 10 s clip).  ``DevicePrefetcher`` keeps ``depth`` batches in flight: each host batch is staged in pinned
memory and copied on a dedicated copy stream while the compute stream works on the previous one; the consumer only
waits on the copy's event."""
import torch


def collate_waveforms(samples, n):
    """[(waveform [m] or [1,m], label, ...)] -> (wav [B,1,n] float32, labels): right zero-pad / truncate to n samples
    (the reference's RandomCrop pads short clips the same way, transforms/common.py:69-72)."""
    B = len(samples)
    wav = torch.zeros(B, 1, n, dtype=torch.float32)
    labels = []
    for b, s in enumerate(samples):
        w = s[0].reshape(-1)
        m = min(n, w.numel())
        wav[b, 0, :m] = w[:m]
        labels.append(s[1])
    try:
        labels = torch.stack([torch.as_tensor(l) for l in labels])
    except Exception:  # noqa: BLE001 - ragged or non-tensor labels stay a list
        pass
    return wav, labels


class DevicePrefetcher:
    """iterate device-resident batches from an iterable of host batches (a tensor or a tuple / list of tensors;
    non-tensor members pass through).  Buffers are reused: steady state allocates nothing."""

    def __init__(self, source, device, depth=2):
        self.source, self.device, self.depth = source, torch.device(device), max(2, depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict(pinned=None, dev=None, ev=None, meta=None) for _ in range(self.depth)]
        self.h2d_bytes = 0

    def _stage(self, slot, batch):
        single = torch.is_tensor(batch)
        parts = [batch] if single else list(batch)
        if slot["dev"] is None or len(slot["dev"]) != len(parts):
            slot["dev"], slot["pinned"] = [None] * len(parts), [None] * len(parts)
        # the previous consumer of this slot's device buffers ran on the compute stream `depth` batches ago
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        nbytes = 0
        with torch.cuda.stream(self.stream):
            for i, t in enumerate(parts):
                if not torch.is_tensor(t):
                    slot["dev"][i] = t
                    continue
                if not t.is_pinned():
                    if slot["pinned"][i] is None or slot["pinned"][i].shape != t.shape or slot["pinned"][i].dtype != t.dtype:
                        slot["pinned"][i] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    slot["pinned"][i].copy_(t)
                    t = slot["pinned"][i]
                d = slot["dev"][i]
                if not torch.is_tensor(d) or d.shape != t.shape or d.dtype != t.dtype:
                    d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                    slot["dev"][i] = d
                d.copy_(t, non_blocking=True)
                nbytes += t.numel() * t.element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        slot["ev"], slot["meta"] = ev, single
        self.h2d_bytes = nbytes

    def __iter__(self):
        it = iter(self.source)
        pending = []
        k = 0
        try:
            for _ in range(self.depth - 1):
                self._stage(self.slots[k % self.depth], next(it))
                pending.append(k % self.depth)
                k += 1
        except StopIteration:
            it = None
        while pending:
            slot = self.slots[pending.pop(0)]
            torch.cuda.current_stream(self.device).wait_event(slot["ev"])
            if it is not None:
                try:
                    self._stage(self.slots[k % self.depth], next(it))
                    pending.append(k % self.depth)
                    k += 1
                except StopIteration:
                    it = None
            yield slot["dev"][0] if slot["meta"] else tuple(slot["dev"])
