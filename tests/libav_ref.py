"""Independent MP3 decoder for the tests: the `mp3float` decoder of the libavcodec that ships inside the
opencv-python-headless wheel of this image, driven through ctypes (packet in, planar f32 out).  Test infrastructure
only — the product never loads it.  `available()` is False where the wheel is absent; the tests then skip."""
import ctypes as C
import glob
import os

_state = {}


def _load():
    if "lib" in _state:
        return _state["lib"]
    lib = util = None
    try:
        import cv2
        d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
        for n in ["libdrm", "libcrypto", "libssl", "libvpx", "libaom", "libavutil", "libswresample", "libavcodec"]:
            hits = glob.glob(os.path.join(d, n + "-*"))
            if not hits:
                continue
            try:
                h = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
            except OSError:
                continue
            if n == "libavcodec":
                lib = h
            if n == "libavutil":
                util = h
    except Exception:
        lib = None
    if lib is not None and util is not None:
        lib.avcodec_find_decoder_by_name.restype = C.c_void_p
        lib.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
        lib.avcodec_alloc_context3.restype = C.c_void_p
        lib.avcodec_alloc_context3.argtypes = [C.c_void_p]
        lib.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.av_packet_alloc.restype = C.c_void_p
        lib.av_new_packet.argtypes = [C.c_void_p, C.c_int]
        lib.av_packet_unref.argtypes = [C.c_void_p]
        lib.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
        lib.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
        util.av_frame_alloc.restype = C.c_void_p
        if not lib.avcodec_find_decoder_by_name(b"mp3float"):
            lib = None
    else:
        lib = None
    _state["lib"], _state["util"] = lib, util
    return lib


def available():
    return _load() is not None


def decode_frames(frames, channels):
    """frames: list of bytes, one MPEG audio frame each -> list of `channels` float32 arrays (planar)."""
    import numpy as np
    lib = _load()
    util = _state["util"]
    codec = lib.avcodec_find_decoder_by_name(b"mp3float")
    ctx = lib.avcodec_alloc_context3(codec)
    assert lib.avcodec_open2(ctx, codec, None) == 0
    pkt, frm = lib.av_packet_alloc(), util.av_frame_alloc()
    out = [[] for _ in range(channels)]
    for fb in frames:
        assert lib.av_new_packet(pkt, len(fb)) == 0
        data = C.c_void_p.from_address(pkt + 24).value          # AVPacket: buf, pts, dts, data, size
        C.memmove(data, fb, len(fb))
        rc = lib.avcodec_send_packet(ctx, pkt)
        lib.av_packet_unref(pkt)
        assert rc == 0, rc
        while lib.avcodec_receive_frame(ctx, frm) == 0:
            n = C.c_int.from_address(frm + 112).value           # AVFrame: data[8], linesize[8], extended_data, width, height, nb_samples
            for c in range(channels):
                p = C.c_void_p.from_address(frm + 8 * c).value
                out[c].append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (n,)).copy())
    return [np.concatenate(o) if o else np.zeros(0, np.float32) for o in out]
