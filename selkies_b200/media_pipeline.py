"""Mirror of the reference's media-pipeline boundary (src/selkies/media_pipeline.py) for the video path.

`MediaPipeline` is the abstract class the rest of selkies programs against (media_pipeline.py:41-80);
`MediaPipelineB200` is the drop-in counterpart of `MediaPipelinePixel` (media_pipeline.py:82-429): same
constructor arguments, same ten methods, same callbacks (`produce_data(buf, pts, kind)`,
`send_data_channel_message(msg)`), same guard/ignore rules — but `capture_module` is
`selkies_b200.pixelflux_compat.ScreenCapture`, i.e. the CUDA pipeline behind libb2video.

Audio (pcmflux, media_pipeline.py:334-395) is outside this tier: `set_audio_bitrate` is accepted and ignored.
"""
from __future__ import annotations

import asyncio
import logging
from abc import ABCMeta, abstractmethod
from enum import Enum

from .pixelflux_compat import CaptureSettings, ScreenCapture

logger = logging.getLogger("media_pipeline_b200")
logger.setLevel(logging.INFO)


class RateControlMode(str, Enum):          # media_pipeline.py:34-36
    CBR = "cbr"
    CRF = "crf"


class MediaPipelineError(Exception):        # media_pipeline.py:38-39
    pass


class MediaPipeline(metaclass=ABCMeta):     # media_pipeline.py:41-80
    @abstractmethod
    def start_media_pipeline(self): ...

    @abstractmethod
    def stop_media_pipeline(self): ...

    @abstractmethod
    def is_media_pipeline_running(self) -> bool: ...

    @abstractmethod
    async def set_pointer_visible(self, visible: bool): ...

    @abstractmethod
    async def set_framerate(self, framerate: int): ...

    @abstractmethod
    async def set_video_bitrate(self, bitrate: int): ...

    @abstractmethod
    async def set_audio_bitrate(self, bitrate: int): ...

    @abstractmethod
    async def dynamic_idr_frame(self): ...

    @abstractmethod
    async def update_rate_control_mode(self, mode: RateControlMode): ...

    @abstractmethod
    async def set_crf(self, crf: int): ...


class MediaPipelineB200(MediaPipeline):
    def __init__(self, async_event_loop: asyncio.AbstractEventLoop, encoder_rtc: str = "x264enc", framerate: int = 30,
                 video_bitrate: int = 8, audio_bitrate: int = 128000, width: int = 1920, height: int = 1080,
                 audio_channels: int = 2, audio_enabled: bool = False, audio_device_name="output.monitor",
                 crf: int = 23, rc_mode: RateControlMode = RateControlMode.CBR, gpu_id: int = 0, frame_source=None):
        self.async_event_loop = async_event_loop
        self.audio_channels = audio_channels
        self.encoder_rtc = encoder_rtc
        self.framerate = framerate
        self.video_bitrate = video_bitrate          # Mbps, as in the reference
        self.rc_mode = rc_mode
        self.h264_crf = crf
        self.audio_bitrate = audio_bitrate
        self.last_resize_success = True
        self.width = width
        self.height = height
        self.audio_enabled = audio_enabled
        self.audio_device_name = audio_device_name
        self.capture_cursor = False
        self.gpu_id = gpu_id
        self.frame_source = frame_source
        self.produce_data = lambda buf, pts, kind: logger.warning("unhandled produce_data")
        self.send_data_channel_message = lambda msg: logger.warning("unhandled send_data_channel_message")
        self.capture_module = None
        self._is_screen_capturing = False
        self._running = False
        self.async_lock = asyncio.Lock()

    # ---- setters: same guards as media_pipeline.py:123-249 -------------------------------------------
    async def set_pointer_visible(self, visible: bool):
        if not self._is_screen_capturing or self.capture_module is None:
            return
        if self.capture_cursor == visible:
            return
        self.capture_cursor = visible
        await self.restart_screen_capture()

    async def update_rate_control_mode(self, mode: RateControlMode):
        if not self._is_screen_capturing or self.capture_module is None:
            return
        if mode == self.rc_mode:
            return
        if mode not in [RateControlMode.CBR, RateControlMode.CRF]:
            logger.error(f"Invalid rate control mode: {mode}")
            return
        self.rc_mode = mode
        try:
            await self.restart_screen_capture()
        except Exception as e:
            logger.info(f"Error updating rate control mode {e}", exc_info=True)

    async def set_crf(self, new_crf: int):
        if not self._is_screen_capturing or self.capture_module is None:
            return
        if self.rc_mode != RateControlMode.CRF or self.h264_crf == new_crf:
            return
        self.h264_crf = new_crf
        try:
            await self.restart_screen_capture()
        except Exception as e:
            logger.info(f"Error updating CRF {e}", exc_info=True)

    async def set_video_bitrate(self, new_bitrate: int):
        """:new_bitrate: Mbps; forwarded to the encoder as kbps (media_pipeline.py:183-201)."""
        if not self._is_screen_capturing or self.capture_module is None:
            return
        if self.rc_mode == RateControlMode.CRF or new_bitrate <= 0 or self.video_bitrate == new_bitrate:
            return
        try:
            await self.async_event_loop.run_in_executor(None, self.capture_module.update_video_bitrate, new_bitrate * 1000)
            self.video_bitrate = new_bitrate
        except AttributeError:
            logger.error("Video capture module does not support video bitrate updation")
        except Exception as e:
            logger.info(f"Error updating video bitrate {e}", exc_info=True)

    async def set_audio_bitrate(self, new_bitrate: int):
        return                                  # audio is not part of the video hot path

    async def set_framerate(self, framerate: int):
        async with self.async_lock:
            if not self._is_screen_capturing:
                return
            if framerate <= 0 or self.framerate == framerate:
                return
            self.framerate = framerate          # also changes the pts step in the callback below
            await self.async_event_loop.run_in_executor(None, self.capture_module.update_framerate, float(self.framerate))

    async def dynamic_idr_frame(self):
        if not self._is_screen_capturing or self.capture_module is None:
            return
        try:
            await self.async_event_loop.run_in_executor(None, self.capture_module.request_idr_frame)
        except AttributeError:
            logger.error("ScreenCapture module does not support IDR frame request")
        except Exception as e:
            logger.error(f"Error requesting IDR frame: {e}", exc_info=True)

    async def set_resolution(self, width: int, height: int):
        """The reference has no such method: WebRTCApp.on_resize_handler writes width/height and pixelflux follows
        (webrtc_mode.py:383-426, media_pipeline.py:261).  Here the follow-up is explicit: new SPS/PPS + IDR."""
        width -= width & 1
        height -= height & 1                     # server rounds down to even (webrtc_mode.py:397-402)
        if width < 16 or height < 16 or width > 7680 or height > 4320:
            self.last_resize_success = False
            return
        self.width, self.height = width, height
        self.last_resize_success = True
        if self._is_screen_capturing and self.capture_module is not None:
            await self.async_event_loop.run_in_executor(None, self.capture_module.update_resolution, width, height)

    # ---- capture start/stop: media_pipeline.py:251-332 ---------------------------------------------------
    def generate_capture_settings(self):
        cs = CaptureSettings()
        cs.capture_width = self.width
        cs.capture_height = self.height
        cs.capture_x = 0
        cs.capture_y = 0
        cs.target_fps = float(self.framerate)
        cs.capture_cursor = self.capture_cursor
        cs.output_mode = 1
        cs.auto_adjust_screen_capture_size = True
        cs.gpu_id = self.gpu_id
        if self.encoder_rtc in ["nvh264enc", "x264enc"]:
            cs.h264_streaming_mode = True
            cs.h264_fullframe = True
            cs.h264_crf = self.h264_crf
            cs.h264_cbr_mode = self.rc_mode == RateControlMode.CBR
            cs.h264_bitrate_kbps = self.video_bitrate * 1000
            cs.vaapi_render_node_index = -1
            if self.encoder_rtc == "x264enc":
                cs.use_cpu = True                # accepted for compatibility; the encode still runs on the GPU
        return cs

    async def start_screen_capture(self):
        if self._is_screen_capturing:
            return
        settings = self.generate_capture_settings()

        def screen_capture_callback(result_ptr, _):
            if not result_ptr:
                return
            try:
                result = result_ptr.contents
                if result.size > 0:
                    data_bytes = bytes(result.data[10:result.size])       # strip the 10-byte stripe header
                    pts_step = 90000 // self.framerate
                    pts = result.frame_id * pts_step
                    asyncio.run_coroutine_threadsafe(self.produce_data(data_bytes, pts, "video"), self.async_event_loop)
            except Exception as e:
                logger.error(f"Error in capture callback: {e}", exc_info=False)

        try:
            self.capture_module = ScreenCapture(self.frame_source)
            await self.async_event_loop.run_in_executor(None, self.capture_module.start_capture, settings, screen_capture_callback)
            self._is_screen_capturing = True
        except Exception as e:
            logger.error(f"Failed to start screen capture: {e}", exc_info=True)
            self.capture_module = None
            self._is_screen_capturing = False

    async def stop_screen_capture(self):
        if not self._is_screen_capturing or self.capture_module is None:
            return
        try:
            await self.async_event_loop.run_in_executor(None, self.capture_module.stop_capture)
        except Exception as e:
            logger.error(f"Error stopping screen capture: {e}", exc_info=True)
        self.capture_module = None
        self._is_screen_capturing = False

    async def restart_screen_capture(self):
        if not self._is_screen_capturing:
            return
        async with self.async_lock:
            try:
                await self.stop_screen_capture()
                await self.start_screen_capture()
            except Exception as e:
                logger.error(f"Error restarting screen capture: {e}")

    async def start_media_pipeline(self):
        async with self.async_lock:
            if self._running:
                return
            try:
                await self.start_screen_capture()
                self._running = self._is_screen_capturing
            except Exception as e:
                logger.error(f"Error starting media pipelines: {e}", exc_info=True)

    async def stop_media_pipeline(self):
        async with self.async_lock:
            if not self._running:
                return
            try:
                await self.stop_screen_capture()
                self._running = False
            except Exception as e:
                logger.error(f"Error stopping media pipelines: {e}", exc_info=True)

    def is_media_pipeline_running(self):
        return self._running
