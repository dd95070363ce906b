"""Asynchronous checkpoints in the reference's format: `torch.save(unet.state_dict(),
path)` as trainer.py:152-233 writes them -- the same 46 keys, shapes and fp32 dtype -- without stalling the training
stream.  The flat parameter arena is copied to a pinned host buffer on a side stream (one 124 MB D2H instead of 46
small ones) and a worker thread serialises it; `wait()` joins the writers (called at the end of training and before a
file is overwritten)."""
import os
import threading

import torch


class AsyncCheckpointer:
    def __init__(self):
        self.stream = None
        self.pinned = []          # reusable pinned buffers (one per save in flight)
        self.threads = []

    def save(self, unet, path):
        eng = getattr(unet, "engine", None)
        st = eng.store if eng is not None else None
        if st is None or st.p_arena is None or not st.p_arena.is_cuda:
            torch.save(unet.state_dict(), path)          # nothing on a device yet: the plain route
            return
        st.adopt(st.p_arena.device)                       # parameters are views of the arena
        dev = st.p_arena.device
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        self.wait(path)                                   # never two writers on one file
        host = self.pinned.pop() if self.pinned else torch.empty(st.total, dtype=torch.float32, pin_memory=True)
        done = torch.cuda.Event()
        self.stream.wait_stream(torch.cuda.current_stream(dev))     # after the optimizer step that produced these weights
        with torch.cuda.stream(self.stream):
            host.copy_(st.p_arena, non_blocking=True)
            done.record(self.stream)
        # the training stream must not overwrite the arena before the copy has read it
        torch.cuda.current_stream(dev).wait_event(done)
        layout = [(name, st.p_off[name], st.numel[name], tuple(p.shape)) for name, p in st.params().items()]

        def work():
            # a failure (disk full, bad path) is kept and re-raised by wait(): the reference's synchronous torch.save
            # would have raised at this point of the training loop
            try:
                done.synchronize()
                sd = {name: host[off:off + n].view(shape).clone() for name, off, n, shape in layout}
                tmp = path + ".tmp"
                torch.save(sd, tmp)
                os.replace(tmp, path)                     # the file appears atomically, complete
            except BaseException as e:  # noqa: BLE001
                t.error = e
            finally:
                self.pinned.append(host)

        t = threading.Thread(target=work, name=f"ckpt:{path}", daemon=False)
        t.path = path
        t.error = None
        t.start()
        self.threads.append(t)

    def wait(self, path=None):
        """Join the writers (of `path` only, if given)."""
        keep, failed = [], None
        for t in self.threads:
            if path is None or t.path == path:
                t.join()
                failed = failed or t.error
            else:
                keep.append(t)
        self.threads = keep
        if failed is not None:
            raise RuntimeError(f"asynchronous checkpoint failed: {failed!r}") from failed
