"""The slice of the kubelet device-plugin gRPC surface that the best-fit path sits behind, as the
reference serves it (pluginapi.DevicePluginServer, k8s.io/kubelet v0.20.2, package `v1beta1`):

  GetDevicePluginOptions   pkg/plugins/base.go:72-76 (and the Register options, :155-157): the
                           reference answers PreStartRequired only; with the path wired in it also
                           advertises GetPreferredAllocationAvailable, or kubelet never asks.
  GetPreferredAllocation   pkg/plugins/base.go:94-96: an empty stub in the reference; here every
                           ContainerPreferredAllocationRequest goes through the C ABI
                           (egpu_preferred_allocation -> CUDA best-fit) and comes back as deviceIDs.

This is the host-side mirror the INTEGRATION.md patch describes for the Go handlers, kept in Python so
that a fake kubelet (grpcio) can drive the path end to end without a Go toolchain; it is not a
re-implementation of the agent (ListAndWatch, Allocate, PreStartContainer, registration stay in Go).
Messages are built from a descriptor written out here - field names and numbers as in
vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto:28-33,48,133-150 - because neither
protoc nor grpc_tools is available in this image.  The device choice is made by the CUDA library only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

SERVICE = "v1beta1.DevicePlugin"
_T = None


def messages():
    """{name: message class} for Empty, DevicePluginOptions, PreferredAllocationRequest/Response and
    their per-container messages."""
    global _T
    if _T is not None:
        return _T
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="egpu/deviceplugin_v1beta1_subset.proto", package="v1beta1", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, number, ftype, label, tname in fields:
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if tname:
                f.type_name = ".v1beta1." + tname
    REP, OPT = F.LABEL_REPEATED, F.LABEL_OPTIONAL
    msg("Empty", [])
    msg("DevicePluginOptions", [("pre_start_required", 1, F.TYPE_BOOL, OPT, None),
                                ("get_preferred_allocation_available", 2, F.TYPE_BOOL, OPT, None)])
    msg("ContainerPreferredAllocationRequest", [("available_deviceIDs", 1, F.TYPE_STRING, REP, None),
                                                ("must_include_deviceIDs", 2, F.TYPE_STRING, REP, None),
                                                ("allocation_size", 3, F.TYPE_INT32, OPT, None)])
    msg("PreferredAllocationRequest", [("container_requests", 1, F.TYPE_MESSAGE, REP, "ContainerPreferredAllocationRequest")])
    msg("ContainerPreferredAllocationResponse", [("deviceIDs", 1, F.TYPE_STRING, REP, None)])
    msg("PreferredAllocationResponse", [("container_responses", 1, F.TYPE_MESSAGE, REP, "ContainerPreferredAllocationResponse")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    _T = {n: message_factory.GetMessageClass(pool.FindMessageTypeByName("v1beta1." + n))
          for n in ("Empty", "DevicePluginOptions", "ContainerPreferredAllocationRequest", "PreferredAllocationRequest",
                    "ContainerPreferredAllocationResponse", "PreferredAllocationResponse")}
    return _T


class BestFitDevicePlugin:
    """GetDevicePluginOptions + GetPreferredAllocation of one plugin (gpu-core or gpu-memory).

    ctx_handle / lib: the egpu_ctx* and the library that exports egpu_preferred_allocation
    (the product library; the CPU tests pass a harness build of csrc/egpu_plugin.cc)."""

    def __init__(self, ctx_handle, resource: int, lib=None):
        self._h = ctx_handle
        self._resource = int(resource)
        self._lib = lib if lib is not None else L.load()
        self._lib.egpu_preferred_allocation.restype = C.c_int

    # pkg/plugins/base.go:72-76 + the one added option
    def GetDevicePluginOptions(self, request, context):
        return messages()["DevicePluginOptions"](pre_start_required=True, get_preferred_allocation_available=True)

    # pkg/plugins/base.go:94-96
    def GetPreferredAllocation(self, request, context):
        T = messages()
        resp = T["PreferredAllocationResponse"]()
        for creq in request.container_requests:
            av = [s.encode() for s in creq.available_deviceIDs]
            mu = [s.encode() for s in creq.must_include_deviceIDs]
            size = int(creq.allocation_size)
            a_arr = (C.c_char_p * max(1, len(av)))(*av)
            m_arr = (C.c_char_p * max(1, len(mu)))(*mu)
            pos = np.full(max(1, size), -1, dtype=np.int32)
            gpu = C.c_int32(-1)
            rc = self._lib.egpu_preferred_allocation(self._h, a_arr, C.c_int64(len(av)), m_arr, C.c_int64(len(mu)), C.c_int32(size),
                                                     C.c_int(self._resource), C.c_void_p(pos.ctypes.data), C.byref(gpu))
            cresp = resp.container_responses.add()
            if rc == L.OK:
                cresp.deviceIDs.extend(creq.available_deviceIDs[p] for p in pos[:size])
            elif rc == L.ERR_UNSAT:
                pass  # empty preference: kubelet falls back to its own choice for this container
            else:     # a Go handler returns (nil, err) -> gRPC status Unknown (pkg/plugins/gpushare.go:41-43 style)
                import grpc
                context.abort(grpc.StatusCode.UNKNOWN, f"egpu_preferred_allocation: code {rc}")
        return resp


def add_to_server(plugin: BestFitDevicePlugin, server):
    """Registers the two unary RPCs under the service name kubelet dials (v1beta1.DevicePlugin)."""
    import grpc
    T = messages()
    handlers = {
        "GetDevicePluginOptions": grpc.unary_unary_rpc_method_handler(
            plugin.GetDevicePluginOptions, request_deserializer=T["Empty"].FromString,
            response_serializer=T["DevicePluginOptions"].SerializeToString),
        "GetPreferredAllocation": grpc.unary_unary_rpc_method_handler(
            plugin.GetPreferredAllocation, request_deserializer=T["PreferredAllocationRequest"].FromString,
            response_serializer=T["PreferredAllocationResponse"].SerializeToString),
    }
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE, handlers),))


class KubeletStub:
    """What kubelet's device manager does with a plugin endpoint, for the two RPCs above."""

    def __init__(self, channel):
        T = messages()
        self.get_options = channel.unary_unary(f"/{SERVICE}/GetDevicePluginOptions", request_serializer=T["Empty"].SerializeToString,
                                               response_deserializer=T["DevicePluginOptions"].FromString)
        self.get_preferred = channel.unary_unary(f"/{SERVICE}/GetPreferredAllocation",
                                                 request_serializer=T["PreferredAllocationRequest"].SerializeToString,
                                                 response_deserializer=T["PreferredAllocationResponse"].FromString)
